// coupled.cu -- translation unit of the cooperative fallback kernel of the coupled pass (coupled.cuh)
#include "coupled.cuh"

cudaError_t coupled_plan(int sms, int *simple_blocks, size_t *simple_smem) {
    cudaError_t e;
    *simple_smem = sizeof(OrgS<MAXC>) * SIMPLE_WARPS;
    if ((e = cudaFuncSetAttribute(k_coupled_simple, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*simple_smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_coupled_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*simple_smem)) != cudaSuccess) return e;
    int ps = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ps, k_coupled_simple, SIMPLE_WARPS * 32, *simple_smem)) != cudaSuccess) return e;
    *simple_blocks = ps * sms;
    return cudaSuccess;
}

cudaError_t coupled_launch(const SapioWorld &w, const Workspace &W, const StepK &K, int blocks, size_t smem, cudaStream_t st) {
    // both forms are launched; the planner's count of organisms in oversize components (FL_BIG_NORG) decides on the device which one works
    k_coupled_cluster<<<CLUSTER_CTAS, SIMPLE_WARPS * 32, smem, st>>>(w, W, K);
    SapioWorld wc = w; Workspace Wc = W; StepK Kc = K;
    void *args[] = { (void *)&wc, (void *)&Wc, (void *)&Kc };
    return cudaLaunchCooperativeKernel((void *)k_coupled_simple, dim3(blocks), dim3(SIMPLE_WARPS * 32), args, smem, st);
}
