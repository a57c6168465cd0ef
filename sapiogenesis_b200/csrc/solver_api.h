// solver_api.h -- host entry points of the solver translation units (org8.cu .. org64.cu, coupled.cu): each size class of the
// organism solver and the coupled kernels are compiled on their own (the templates dominate the build time).
#pragma once
#include "steptypes.cuh"

#define ORG_CLASSES 8
__host__ __device__ inline int org_class_of(int nc) { return nc <= 8 ? 0 : (nc - 1) >> 3; }      // rows 1..8 -> 0, 9..16 -> 1, ...

typedef cudaError_t (*OrgPlanFn)(int sms, int *blocks, size_t *smem);
typedef void (*OrgLaunchFn)(const SapioWorld &w, const Workspace &W, const StepK &K, int cls, int blocks, size_t smem, cudaStream_t ss);
typedef void (*OrgInfoFn)(int *out);
#define ORG_DECL(CM) cudaError_t org_plan_##CM(int sms, int *blocks, size_t *smem); void org_info_##CM(int *out); \
    void org_launch_##CM(const SapioWorld &w, const Workspace &W, const StepK &K, int cls, int blocks, size_t smem, cudaStream_t ss);
ORG_DECL(8) ORG_DECL(16) ORG_DECL(24) ORG_DECL(32) ORG_DECL(40) ORG_DECL(48) ORG_DECL(56) ORG_DECL(64)
#undef ORG_DECL
static const OrgPlanFn ORG_PLAN[ORG_CLASSES] = {org_plan_8, org_plan_16, org_plan_24, org_plan_32, org_plan_40, org_plan_48, org_plan_56, org_plan_64};
static const OrgInfoFn ORG_INFO[ORG_CLASSES] = {org_info_8, org_info_16, org_info_24, org_info_32, org_info_40, org_info_48, org_info_56, org_info_64};
static const OrgLaunchFn ORG_LAUNCH_FN[ORG_CLASSES] = {org_launch_8, org_launch_16, org_launch_24, org_launch_32, org_launch_40, org_launch_48, org_launch_56, org_launch_64};

// coupled.cu: the cooperative fallback; component.cu: connected components of the cross-contact graph, one per CTA
cudaError_t coupled_plan(int sms, int *simple_blocks, size_t *simple_smem);
cudaError_t coupled_launch(const SapioWorld &w, const Workspace &W, const StepK &K, int blocks, size_t smem, cudaStream_t st);
#define COMP_TIERS_N 7
cudaError_t component_plan(int sms, int *blocks);
void component_plan_launch(const SapioWorld &w, const Workspace &W, int force_fallback, int grid, cudaStream_t st);
void component_launch(const SapioWorld &w, const Workspace &W, const StepK &K, int tier, int blocks, cudaStream_t st);
