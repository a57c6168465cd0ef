// org_tu.cuh -- one size class of the organism solver per translation unit
#pragma once
#include "orgsolve.cuh"
#define ORG_TU(CM) \
cudaError_t org_plan_##CM(int sms, int *blocks, size_t *smem_out) { \
    size_t smem = (size_t)OrgCfg<CM>::WARP_BYTES * OrgCfg<CM>::WARPS; \
    cudaError_t e = cudaFuncSetAttribute(k_org_solve<CM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e; \
    e = cudaFuncSetAttribute(k_org_solve<CM>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);   /* two CTAs of 111 KB need the whole 228 KB */ \
    if (e != cudaSuccess) return e; \
    int per_sm = 0; \
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_org_solve<CM>, OrgCfg<CM>::WARPS * 32, smem); \
    if (e != cudaSuccess) return e; \
    *smem_out = smem; *blocks = (per_sm > 0 ? per_sm : 1) * sms;     /* persistent: every SM full, work pulled dynamically */ \
    return cudaSuccess; \
} \
void org_info_##CM(int *out) { out[0] = OrgCfg<CM>::G; out[1] = OrgCfg<CM>::NG; out[2] = OrgCfg<CM>::WARPS; out[3] = OrgCfg<CM>::WARP_BYTES; } \
void org_launch_##CM(const SapioWorld &w, const Workspace &W, const StepK &K, int cls, int blocks, size_t smem, cudaStream_t ss) { \
    k_org_solve<CM><<<blocks, OrgCfg<CM>::WARPS * 32, smem, ss>>>(w, W, K, cls); \
}
