// solver_api.h -- host entry points of the solver translation units (org8.cu .. org64.cu, coupled.cu): each size class of the
// organism solver and the coupled kernels are compiled on their own (the templates dominate the build time).
#pragma once
#include "steptypes.cuh"

#define ORG_CLASSES 8
__host__ __device__ inline int org_class_of(int nc) { return nc <= 8 ? 0 : (nc - 1) >> 3; }      // rows 1..8 -> 0, 9..16 -> 1, ...
#define COUPLED_PACK_FROM 2400            /* coupled organisms (last tick) from which the packed form is launched: ~4 per warp of the simple one */

typedef cudaError_t (*OrgPlanFn)(int sms, int *blocks, size_t *smem);
typedef void (*OrgLaunchFn)(const SapioWorld &w, const Workspace &W, const StepK &K, int cls, int blocks, size_t smem, cudaStream_t ss);
#define ORG_DECL(CM) cudaError_t org_plan_##CM(int sms, int *blocks, size_t *smem); \
    void org_launch_##CM(const SapioWorld &w, const Workspace &W, const StepK &K, int cls, int blocks, size_t smem, cudaStream_t ss);
ORG_DECL(8) ORG_DECL(16) ORG_DECL(24) ORG_DECL(32) ORG_DECL(40) ORG_DECL(48) ORG_DECL(56) ORG_DECL(64)
#undef ORG_DECL
static const OrgPlanFn ORG_PLAN[ORG_CLASSES] = {org_plan_8, org_plan_16, org_plan_24, org_plan_32, org_plan_40, org_plan_48, org_plan_56, org_plan_64};
static const OrgLaunchFn ORG_LAUNCH_FN[ORG_CLASSES] = {org_launch_8, org_launch_16, org_launch_24, org_launch_32, org_launch_40, org_launch_48, org_launch_56, org_launch_64};

cudaError_t coupled_plan(int sms, int *coop_blocks, size_t *coop_smem, int *simple_blocks, size_t *simple_smem);
cudaError_t coupled_launch(const SapioWorld &w, const Workspace &W, const StepK &K, bool packed, int blocks, size_t smem, cudaStream_t st);
