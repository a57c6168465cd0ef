// coupled.cu -- translation unit of the coupled-organism kernels (coupled.cuh)
#include "coupled.cuh"

cudaError_t coupled_plan(int sms, int *coop_blocks, size_t *coop_smem, int *simple_blocks, size_t *simple_smem) {
    cudaError_t e;
    *simple_smem = sizeof(OrgS<MAXC>) * SIMPLE_WARPS;
    if ((e = cudaFuncSetAttribute(k_coupled_simple, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*simple_smem)) != cudaSuccess) return e;
    int ps = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ps, k_coupled_simple, SIMPLE_WARPS * 32, *simple_smem)) != cudaSuccess) return e;
    *simple_blocks = ps * sms;
    *coop_smem = CP_SLAB * COUPLED_WARPS;
    if ((e = cudaFuncSetAttribute(k_coupled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*coop_smem)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ps, k_coupled, COUPLED_WARPS * 32, *coop_smem)) != cudaSuccess) return e;
    *coop_blocks = ps * sms;
    return cudaSuccess;
}

cudaError_t coupled_launch(const SapioWorld &w, const Workspace &W, const StepK &K, bool packed, int blocks, size_t smem, cudaStream_t st) {
    SapioWorld wc = w; Workspace Wc = W; StepK Kc = K;
    void *args[] = { (void *)&wc, (void *)&Wc, (void *)&Kc };
    if (packed) return cudaLaunchCooperativeKernel((void *)k_coupled, dim3(blocks), dim3(COUPLED_WARPS * 32), args, smem, st);
    return cudaLaunchCooperativeKernel((void *)k_coupled_simple, dim3(blocks), dim3(SIMPLE_WARPS * 32), args, smem, st);
}
