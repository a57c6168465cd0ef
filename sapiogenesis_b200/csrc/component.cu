// component.cu -- translation unit of the component solver (components.cuh)
#include "components.cuh"

cudaError_t component_plan(int sms, int *blocks /* [COMP_TIERS] */) {
    for (int t = 0; t < COMP_TIERS; t++) {
        const CompTier T = comp_tier(t);
        cudaError_t e = cudaSuccess;
        if (t == COMP_TIERS - 1 && (e = cudaFuncSetAttribute(k_component, cudaFuncAttributeMaxDynamicSharedMemorySize, T.smem)) != cudaSuccess) return e;
    }
    {   // the attribute is per function: the largest tier's size admits every launch
        const CompTier T = comp_tier(COMP_TIERS - 1);
        cudaError_t e = cudaFuncSetAttribute(k_component, cudaFuncAttributeMaxDynamicSharedMemorySize, T.smem);
        if (e != cudaSuccess) return e;
    }
    for (int t = 0; t < COMP_TIERS; t++) {
        const CompTier T = comp_tier(t);
        int ps = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ps, k_component, T.warps * 32, (size_t)T.smem);
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
    k_component<<<blocks, T.warps * 32, (size_t)T.smem, st>>>(w, W, K, tier);
}
