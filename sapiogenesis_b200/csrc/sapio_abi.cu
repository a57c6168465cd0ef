// sapio_abi.cu -- the C-ABI of libsapio_b200.so (include/sapio.h). Thin: argument checks, workspace carve-up,
// kernel launches on the caller's stream. No allocation, no torch types, no CPU fallback: without a device every
// compute entry point returns SAPIO_E_NODEVICE.
#include <cstdio>
#include <cstring>
#include <cub/cub.cuh>
#include "common.cuh"
#include "bodies.cuh"
#include "step.cuh"

static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, const char *a = "") {
    snprintf(g_err, sizeof(g_err), fmt, a);
    return code;
}
#define CUDA_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return fail(SAPIO_E_CUDA, #expr ": %s", cudaGetErrorString(e_)); } while (0)
#define LAUNCH_CHECK() CUDA_TRY(cudaGetLastError())

static int g_sms = 0;
static int sm_count() {
    if (!g_sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_sms;
}
// streaming kernels: a whole number of waves over the SMs (148 on B200), 8 CTAs of 256 threads resident per SM
static int stream_grid(long items, int threads = 256) {
    int sms = sm_count() > 0 ? sm_count() : 148;
    long want = (items + threads - 1) / threads;
    long wave = (long)sms * 8;
    if (want <= wave) return (int)(want > 0 ? want : 1);
    return (int)wave;
}

extern "C" {

int sapio_abi_version(void) { return SAPIO_ABI_VERSION; }
const char *sapio_last_error(void) { return g_err; }
int sapio_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

size_t sapio_workspace_bytes(const SapioParams *p) {
    if (!p) return 0;
    Workspace ws;
    return carve_workspace(*p, nullptr, &ws);
}

static int check(const SapioWorld *w, void *ws, size_t ws_bytes, Workspace *out) {
    if (!w) return fail(SAPIO_E_ARG, "null world%s");
    if (sapio_device_count() <= 0) return fail(SAPIO_E_NODEVICE, "no CUDA device: libsapio_b200 has no CPU path%s");
    if (out) {
        size_t need = carve_workspace(w->p, (char *)ws, out);
        if (!ws || ws_bytes < need) return fail(SAPIO_E_WORKSPACE, "workspace too small%s");
    }
    return 0;
}

int sapio_friction(const SapioWorld *w, void *stream) {
    int rc = check(w, nullptr, 0, nullptr); if (rc) return rc;
    long nb = (long)w->p.n_org * MAXC + w->p.n_dead;
    k_friction<<<stream_grid(nb), 256, 0, (cudaStream_t)stream>>>(*w);
    LAUNCH_CHECK();
    return 0;
}

int sapio_post(const SapioWorld *w, void *ws, size_t ws_bytes, void *stream) {
    Workspace W; int rc = check(w, ws, ws_bytes, &W); if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(W.pop, 0, sizeof(int), st));
    k_post_count<<<stream_grid(w->p.n_org), 256, 0, st>>>(*w, W.pop);
    k_post_finish<<<1, 1, 0, st>>>(*w, W.pop, 1, 0);
    LAUNCH_CHECK();
    return 0;
}

int sapio_space_step(const SapioWorld *w, void *ws, size_t ws_bytes, void *stream) {
    Workspace W; int rc = check(w, ws, ws_bytes, &W); if (rc) return rc;
    return launch_space_step(*w, W, (cudaStream_t)stream, g_err, sizeof(g_err));
}

int sapio_tick(const SapioWorld *w, void *ws, size_t ws_bytes, uint32_t phases, int n_ticks, void *stream) {
    Workspace W; int rc = check(w, ws, ws_bytes, &W); if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    long nb = (long)w->p.n_org * MAXC + w->p.n_dead;
    for (int t = 0; t < n_ticks; t++) {
        if (phases & SAPIO_PH_STEP) { rc = launch_space_step(*w, W, st, g_err, sizeof(g_err)); if (rc) return rc; }
        if (phases & SAPIO_PH_FRICTION) k_friction<<<stream_grid(nb), 256, 0, st>>>(*w);
        if (phases & SAPIO_PH_POST) {
            CUDA_TRY(cudaMemsetAsync(W.pop, 0, sizeof(int), st));
            k_post_count<<<stream_grid(w->p.n_org), 256, 0, st>>>(*w, W.pop);
        }
        k_post_finish<<<1, 1, 0, st>>>(*w, W.pop, (phases & SAPIO_PH_POST) ? 1 : 0, 1);
        LAUNCH_CHECK();
    }
    return 0;
}

}  // extern "C"
