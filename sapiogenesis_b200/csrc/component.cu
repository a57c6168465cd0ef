// component.cu -- translation unit of the component solver (components.cuh)
#include "components.cuh"

typedef void (*CompKernel)(const SapioWorld, Workspace, StepK, int);
static CompKernel comp_kernel(int warps) {
    switch (warps) { case 1: return k_component<1>; case 2: return k_component<2>; case 4: return k_component<4>; default: return k_component<16>; }
}

cudaError_t component_plan(int sms, int *blocks /* [COMP_TIERS] */) {
    for (int t = 0; t < COMP_TIERS; t++) {
        const CompTier T = comp_tier(t);
        CompKernel k = comp_kernel(T.warps);
        // the attribute is per function (instantiation): the largest tier of this CTA size admits every launch
        int need = 0;
        for (int u = 0; u < COMP_TIERS; u++) if (comp_tier(u).warps == T.warps && comp_tier(u).smem > need) need = comp_tier(u).smem;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        int ps = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ps, k, T.warps * 32, (size_t)T.smem);
        if (e != cudaSuccess) return e;
        blocks[t] = (ps > 0 ? ps : 1) * sms;
    }
    return cudaSuccess;
}
void component_plan_launch(const SapioWorld &w, const Workspace &W, int force_fallback, int grid, cudaStream_t st) {
    k_comp_init<<<grid, 256, 0, st>>>(w, W);
    k_comp_union<<<grid, 256, 0, st>>>(w, W);
    k_comp_count<<<grid, 256, 0, st>>>(w, W);
    k_comp_assign<<<grid, 256, 0, st>>>(w, W, force_fallback);
    k_comp_scatter<<<grid, 256, 0, st>>>(w, W);
}

void component_launch(const SapioWorld &w, const Workspace &W, const StepK &K, int tier, int blocks, cudaStream_t st) {
    const CompTier T = comp_tier(tier);
    comp_kernel(T.warps)<<<blocks, T.warps * 32, (size_t)T.smem, st>>>(w, W, K, tier);
}
