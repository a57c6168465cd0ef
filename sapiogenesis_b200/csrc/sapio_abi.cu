// sapio_abi.cu -- the C-ABI of libsapio_b200.so (include/sapio.h). Thin: argument checks, workspace carve-up,
// kernel launches on the caller's stream. No allocation, no torch types, no CPU fallback: without a device every
// compute entry point returns SAPIO_E_NODEVICE.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cub/cub.cuh>
#include "common.cuh"
#include "bodies.cuh"
#include "step.cuh"
#include "rules.cuh"
#include "think.cuh"
#include "train.cuh"
#include "repro.cuh"
#include "frame.cuh"
#include "migrate.cuh"
#include "migrate2.cuh"
#include "halo.cuh"

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
// streaming kernels: a whole number of waves over the SMs (148 on B200), 8 CTAs resident per SM
static int stream_grid(long items, int threads = 256) {
    int sms = sm_count() > 0 ? sm_count() : 148;
    long want = (items + threads - 1) / threads;
    long wave = (long)sms * 8;
    if (want <= wave) return (int)(want > 0 ? want : 1);
    return (int)wave;
}

// scratch shared by the gas scan (one Aff per chunk of organisms) and the reductions (REDUCE_CTAS doubles)
static size_t cub_need(const SapioParams &p) {
    size_t a = ((size_t)p.n_org / GAS_CHUNK + 2) * sizeof(Aff), b = (size_t)REDUCE_CTAS * sizeof(double);
    return (a > b ? a : b) + 256;
}

struct AllWs { Workspace step; RulesWs rules; TrainWs train; ReproWs repro; FrameWs frame, frame2; MigWs mig; HaloWs halo; int train_ctas; int *probe; };

static size_t carve_all(const SapioParams &p, char *base, AllWs *A) {
    size_t off = carve_workspace(p, base, &A->step);
    auto take = [&](size_t bytes) -> char * { char *r = base ? base + off : nullptr; off += (bytes + 255) & ~size_t(255); return r; };
    const size_t NO = (size_t)p.n_org, ND = (size_t)p.n_dead;
    RulesWs &R = A->rules;
    R.gas_map = (Aff *)take(NO * sizeof(Aff)); R.gas_scan = (Aff *)take(NO * sizeof(Aff));
    R.think_list = (int *)take(NO * 4); R.train_list = (int *)take(NO * 4); R.repro_list = (int *)take(NO * 4);
    R.counters = (int *)take(256);
    R.refund_sum = (double *)take(256);
    R.dying_cnt = (int *)take((NO + 1) * 4); R.dying_off = (int *)take((NO + 1) * 4);
    R.free_flag = (int *)take((ND + 1) * 4); R.free_off = (int *)take((ND + 1) * 4); R.free_list = (int *)take(ND * 4);
    R.disp_factor = (double *)take(ND * 8); R.disp_prod = (double *)take(256);
    R.slab_co2_in = (double *)take(256); R.dead_ghost = (uint8_t *)take(ND);
    A->train_ctas = (int)std::min<long>(2 * TRAIN_CTAS_PER_SM * 148, std::max<long>(8, NO / 8));      // two waves of resident CTAs (work counter); ~1/17 of the organisms train per tick
    TrainWs &T = A->train;
    T.per_cta = (size_t)(p.in_cap + 4 * TRAIN_MAXW + 4) * SAPIO_MEMROWS;
    T.buf = (float *)take((size_t)A->train_ctas * T.per_cta * 4);
    ReproWs &Q = A->repro;
    Q.birth_cap = (int)(NO < 4096 ? NO : 4096);
    Q.births = (Birth *)take((size_t)Q.birth_cap * sizeof(Birth));
    Q.req_flag = (int *)take((NO + 1) * 4); Q.req_off = (int *)take((NO + 1) * 4);
    Q.free_flag = (int *)take((NO + 1) * 4); Q.free_off = (int *)take((NO + 1) * 4);
    Q.req_list = (int *)take(NO * 4); Q.free_list = (int *)take(NO * 4);
    Q.accepted = (int *)take((size_t)Q.birth_cap * 4);
    Q.counters = (int *)take(256);
    A->probe = (int *)take(256);
    {
        MigWs &M = A->mig; const size_t R = SAPIO_MIG_MAX_RECORDS;
        M.emig = (int *)take(R * 4); M.counters = (int *)take(256); M.rect = (double *)take(R * 6 * 8); M.mask = (uint32_t *)take(R * 12 * 4); M.upto = (int *)take(R * 4);
        M.slot = (int *)take(R * 4); M.delta = (double *)take(R * 2 * 8); M.acc_rect = (double *)take(R * 4 * 8);
    }
    if (p.slab_mode) {
        HaloWs &H = A->halo;
        for (int sd = 0; sd < 2; sd++) { H.ghost_slot[sd] = (int *)take(NO * 4); H.ghost_dslot[sd] = (int *)take(ND * 4); }
        H.refreshed = (int *)take(NO * 4); H.drefreshed = (int *)take(ND * 4); H.dcount = (int *)take(256); H.dlist = (int *)take(ND * 4);
    }
    FrameWs &F = A->frame;
    F.control = (SapioControl *)take(256); F.header = (SapioFrameHeader *)take(512); F.counts = (int *)take(256);
    F.cells = (float4 *)take(NO * MAXC * sizeof(float4)); F.dead = (float4 *)take(ND * sizeof(float4));
    FrameWs &F2 = A->frame2;                               // second device frame: the pipelined exchange packs k + 1 while k downloads
    F2.control = F.control; F2.header = (SapioFrameHeader *)take(512); F2.counts = (int *)take(256);
    F2.cells = (float4 *)take(NO * MAXC * sizeof(float4)); F2.dead = (float4 *)take(ND * sizeof(float4));
    // cub temp storage is shared by every scan / reduction / sort: one region sized for the largest user
    size_t need = cub_need(p);
    if (need > A->step.cub_bytes) { A->step.cub_tmp = take(need); A->step.cub_bytes = need; }
    return off;
}

static bool g_tables_ready = false;
static int ensure_tables() {
    if (g_tables_ready) return 0;
    double c[360], s[360];
    for (int b = 0; b < 360; b++) { double rad = b * (M_PI / 180.0); c[b] = cos(rad); s[b] = sin(rad); }   // math.radians / math.cos
    CUDA_TRY(cudaMemcpyToSymbol(C_COS, c, sizeof(c)));
    CUDA_TRY(cudaMemcpyToSymbol(C_SIN, s, sizeof(s)));
    g_tables_ready = true;
    return 0;
}

static int check(const SapioWorld *w, void *ws, size_t ws_bytes, AllWs *out) {
    if (!w) return fail(SAPIO_E_ARG, "null world%s");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return fail(SAPIO_E_NODEVICE, "no CUDA device: libsapio_b200 has no CPU path%s"); }
    if (w->p.width_cap > TRAIN_MAXW || w->p.in_cap > 128) return fail(SAPIO_E_ARG, "width_cap / in_cap above the compiled maximum (128)%s");
    if (out) {
        size_t need = carve_all(w->p, (char *)ws, out);
        if (!ws || ws_bytes < need) return fail(SAPIO_E_WORKSPACE, "workspace too small%s");
    }
    return ensure_tables();
}

// ---- stage launchers ----------------------------------------------------------------------------------------------
static int launch_grid_only(const SapioWorld &w, const Workspace &W, cudaStream_t st) {
    const int NC = w.p.n_org * MAXC, NB = NC + w.p.n_dead, NG = w.p.grid_nx * w.p.grid_ny;
    CUDA_TRY(cudaMemsetAsync(W.cell_cnt, 0, (size_t)(NG + 1) * 4, st));
    CUDA_TRY(cudaMemsetAsync(W.flags + FL_NLIVE, 0, sizeof(int), st));
    k_integrate<false><<<stream_grid(NB), 256, 0, st>>>(w, W.gkey, W.gval, W.cell_cnt, W.flags + FL_NLIVE);
    scan_exclusive_i32(W.cell_cnt, W.cell_start, NG + 1, W.scan_parts, st);
    k_grid_scatter<<<stream_grid(NB / 3), 256, 0, st>>>(W, W.flags + FL_NLIVE);
    k_grid_rank<<<stream_grid(NB / 3), 256, 0, st>>>(w, W, W.flags + FL_NLIVE);
    return 0;
}

// k_think pulls organisms from a work counter: one wave of resident CTAs (or fewer, for a short list)
static int think_grid(long organisms) {
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const long want = (organisms + THINK_WARPS - 1) / THINK_WARPS, wave = (long)sms * THINK_MINB;
    return (int)(want < wave ? (want > 0 ? want : 1) : wave);
}

static int launch_rules(const SapioWorld &w, AllWs &A, uint32_t phases, cudaStream_t st, int part = 3) {   // part: 1 begin + gas scan + think, 2 main + finish
    const SapioParams &p = w.p;
    RulesWs &R = A.rules;
    const int org_grid = stream_grid((long)p.n_org * 32, RULES_WARPS * 32);
    if (part & 1) {
    prof_mark(st, PT_RULES_BEGIN);
    CUDA_TRY(cudaMemsetAsync(R.counters, 0, 256, st));
    k_rules_begin<<<org_grid, RULES_WARPS * 32, 0, st>>>(w, R, phases);
    prof_mark(st, PT_GAS_SCAN);
    {
        const int nchunks = (p.n_org + GAS_CHUNK - 1) / GAS_CHUNK;
        if ((size_t)nchunks * sizeof(Aff) > A.step.cub_bytes) return fail(SAPIO_E_ARG, "workspace too small for the gas scan%s");
        Aff *agg = (Aff *)A.step.cub_tmp;
        k_gas_local<<<nchunks, GAS_THREADS, 0, st>>>(R.gas_map, R.gas_scan, agg, p.n_org);
        k_gas_chunks<<<1, 32, 0, st>>>(agg, nchunks);
        if (nchunks > 1) k_gas_apply<<<nchunks - 1, GAS_THREADS, 0, st>>>(R.gas_scan, agg, p.n_org);
        LAUNCHES(nchunks > 1 ? 3 : 2);
    }
    if (phases & SAPIO_PH_THINK) {
        prof_mark(st, PT_GRID);
        if (!(phases & SAPIO_PH_REUSE_GRID)) { int rc = launch_grid_only(w, A.step, st); if (rc) return rc; }
        prof_mark(st, PT_THINK);
        size_t smem = (size_t)THINK_WARPS * (p.in_cap + 2 * (p.width_cap > SAPIO_RNN_MAXH ? p.width_cap : SAPIO_RNN_MAXH)) * sizeof(float);
        static size_t smem_set = 0;
        if (smem > smem_set) { CUDA_TRY(cudaFuncSetAttribute(k_think, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); smem_set = smem; }
        CUDA_TRY(cudaMemsetAsync(R.counters + RC_THINK_TICKET, 0, sizeof(int), st));
        k_think<<<think_grid((long)p.n_org / 4 + 1), THINK_WARPS * 32, smem, st>>>(w, A.step, R);
        LAUNCHES(3);
    }
    }
    if (!(part & 2)) { prof_mark(st, PT_END); LAUNCH_CHECK(); return 0; }
    prof_mark(st, PT_RULES_MAIN);
    k_rules_main<<<org_grid, RULES_WARPS * 32, 0, st>>>(w, R, phases);
    reduce_doubles<false>(w.org_gas_refund, p.n_org, R.refund_sum, (double *)A.step.cub_tmp, st);
    k_rules_finish<<<1, 1, 0, st>>>(w, R);
    LAUNCHES(3);
    prof_mark(st, PT_END);
    LAUNCH_CHECK();
    return 0;
}

static int launch_train(const SapioWorld &w, AllWs &A, cudaStream_t st) {
    const SapioParams &p = w.p;
    int ctas = p.n_org < A.train_ctas ? p.n_org : A.train_ctas;
    prof_mark(st, PT_TRAIN);
    CUDA_TRY(cudaMemsetAsync(A.rules.counters + RC_TRAIN_TICKET, 0, sizeof(int), st));
    k_train<<<ctas, TRAIN_THREADS, 0, st>>>(w, A.rules, A.train);
    LAUNCHES(1);
    prof_mark(st, PT_END);
    LAUNCH_CHECK();
    return 0;
}

static int launch_settle(const SapioWorld &w, AllWs &A, cudaStream_t st) {
    const SapioParams &p = w.p;
    RulesWs &R = A.rules;
    const long NB = (long)p.n_org * MAXC + p.n_dead;
    prof_mark(st, PT_SETTLE);
    k_settle_gather<<<stream_grid(NB), 256, 0, st>>>(w);
    k_settle_count<<<stream_grid((p.n_org > p.n_dead ? p.n_org : p.n_dead) + 1), 256, 0, st>>>(w, R);
    scan_exclusive_i32(R.dying_cnt, R.dying_off, p.n_org + 1, A.step.scan_parts, st);
    scan_exclusive_i32(R.free_flag, R.free_off, p.n_dead + 1, A.step.scan_parts, st);
    k_settle_freelist<<<stream_grid(p.n_dead), 256, 0, st>>>(w, R);
    k_settle_deaths<<<stream_grid(p.n_org), 256, 0, st>>>(w, R);
    k_settle_disperse<<<stream_grid(NB), 256, 0, st>>>(w, R);
    reduce_doubles<true>(R.disp_factor, p.n_dead, R.disp_prod, (double *)A.step.cub_tmp, st);
    k_settle_gas<<<1, 1, 0, st>>>(w, R);
    LAUNCHES(6);
    prof_mark(st, PT_END);
    LAUNCH_CHECK();
    return 0;
}

static int launch_reproduce(const SapioWorld &w, AllWs &A, cudaStream_t st) {
    const SapioParams &p = w.p;
    ReproWs &Q = A.repro;
    prof_mark(st, PT_REPRODUCE);
    for (int sib = 0; sib < p.offspring_amount; sib++) {
        k_repro_flags<<<stream_grid(p.n_org + 1), 256, 0, st>>>(w, Q);
        scan_exclusive_i32(Q.req_flag, Q.req_off, p.n_org + 1, A.step.scan_parts, st);
        scan_exclusive_i32(Q.free_flag, Q.free_off, p.n_org + 1, A.step.scan_parts, st);
        k_repro_lists<<<stream_grid(p.n_org), 256, 0, st>>>(w, Q);
        int blocks = (Q.birth_cap + REPRO_WARPS - 1) / REPRO_WARPS;
        int cap = (sm_count() > 0 ? sm_count() : 148) * 8;
        if (blocks > cap) blocks = cap;
        if (sib == 0) k_org_rects<<<stream_grid(p.n_org), 256, 0, st>>>(w, Q.counters);
        k_repro_prepare<<<blocks, REPRO_WARPS * 32, 0, st>>>(w, Q, sib);
        k_repro_near<<<stream_grid(p.n_org), 256, 0, st>>>(w, Q);
        k_repro_place<<<blocks, REPRO_WARPS * 32, 0, st>>>(w, Q);
        k_repro_resolve<<<1, 32, 0, st>>>(w, Q, sib);
        k_repro_finish<<<blocks, REPRO_WARPS * 32, 0, st>>>(w, Q, sib);
        k_repro_clear<<<stream_grid(p.n_dead), 256, 0, st>>>(w, Q);
        if (sib + 1 < p.offspring_amount) k_repro_next_sibling<<<stream_grid(Q.birth_cap), 256, 0, st>>>(w, Q);
        LAUNCHES(sib == 0 ? 11 : 10);
    }
    prof_mark(st, PT_END);
    LAUNCH_CHECK();
    return 0;
}

static int launch_post(const SapioWorld &w, AllWs &A, cudaStream_t st, int do_post, int advance) {
    prof_mark(st, PT_POST);
    LAUNCHES(do_post ? 2 : 1);
    if (do_post) {
        CUDA_TRY(cudaMemsetAsync(A.step.pop, 0, sizeof(int), st));
        k_post_count<<<stream_grid(w.p.n_org), 256, 0, st>>>(w, A.step.pop);
    }
    k_post_finish<<<1, 1, 0, st>>>(w, A.step.pop, do_post, advance);
    prof_mark(st, PT_END);
    LAUNCH_CHECK();
    return 0;
}

static int launch_tick(const SapioWorld &w, AllWs &A, uint32_t phases, const SapioControl *d_control, cudaStream_t st) {
    int rc;
    const long nb = (long)w.p.n_org * MAXC + w.p.n_dead;
    if (phases & SAPIO_PH_RULES) {
        if ((rc = launch_rules(w, A, phases, st))) return rc;
        if (phases & SAPIO_PH_TRAIN) if ((rc = launch_train(w, A, st))) return rc;
        if ((rc = launch_settle(w, A, st))) return rc;
        if (phases & SAPIO_PH_REPRODUCE) if ((rc = launch_reproduce(w, A, st))) return rc;
    }
    if (phases & SAPIO_PH_STEP) { rc = launch_space_step(w, A.step, st, g_err, sizeof(g_err)); if (rc) return rc; }
    if (phases & SAPIO_PH_FRICTION) { prof_mark(st, PT_FRICTION); k_friction<<<stream_grid(nb), 256, 0, st>>>(w); LAUNCHES(1); }
    if (d_control) { k_events<<<stream_grid(nb), 256, 0, st>>>(w, d_control); k_event_food<<<1, 256, 0, st>>>(w, d_control); LAUNCHES(2); }
    return launch_post(w, A, st, (phases & SAPIO_PH_POST) ? 1 : 0, 1);
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
    AllWs A;
    return carve_all(*p, nullptr, &A);
}

int sapio_friction(const SapioWorld *w, void *stream) {
    int rc = check(w, nullptr, 0, nullptr); if (rc) return rc;
    long nb = (long)w->p.n_org * MAXC + w->p.n_dead;
    k_friction<<<stream_grid(nb), 256, 0, (cudaStream_t)stream>>>(*w);
    LAUNCH_CHECK();
    return 0;
}
int sapio_post(const SapioWorld *w, void *ws, size_t ws_bytes, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    return launch_post(*w, A, (cudaStream_t)stream, 1, 0);
}
int sapio_space_step(const SapioWorld *w, void *ws, size_t ws_bytes, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    return launch_space_step(*w, A.step, (cudaStream_t)stream, g_err, sizeof(g_err));
}
int sapio_rules(const SapioWorld *w, void *ws, size_t ws_bytes, uint32_t phases, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    return launch_rules(*w, A, phases, (cudaStream_t)stream);
}
int sapio_train(const SapioWorld *w, void *ws, size_t ws_bytes, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    return launch_train(*w, A, (cudaStream_t)stream);
}
int sapio_settle(const SapioWorld *w, void *ws, size_t ws_bytes, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    return launch_settle(*w, A, (cudaStream_t)stream);
}
__global__ void k_fill_list(int *list, int *counter, const int32_t *src, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) list[i] = src[i];
    if (i == 0) *counter = n;
}
static int check_list(const SapioWorld *w, const int32_t *d_orgs, int n) {
    if (n < 0 || n > w->p.n_org || (n && !d_orgs)) { snprintf(g_err, sizeof(g_err), "bad organism list (n = %d)", n); return SAPIO_E_ARG; }
    return 0;
}
int sapio_activate(const SapioWorld *w, void *ws, size_t ws_bytes, const int32_t *d_orgs, int n, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    if ((rc = check_list(w, d_orgs, n))) return rc;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const SapioParams &p = w->p;
    k_fill_list<<<(n + 255) / 256, 256, 0, st>>>(A.rules.think_list, A.rules.counters + RC_NTHINK, d_orgs, n);
    if ((rc = launch_grid_only(*w, A.step, st))) return rc;
    size_t smem = (size_t)THINK_WARPS * (p.in_cap + 2 * (p.width_cap > SAPIO_RNN_MAXH ? p.width_cap : SAPIO_RNN_MAXH)) * sizeof(float);
    CUDA_TRY(cudaFuncSetAttribute(k_think, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaMemsetAsync(A.rules.counters + RC_THINK_TICKET, 0, sizeof(int), st));
    k_think<<<think_grid(n), THINK_WARPS * 32, smem, st>>>(*w, A.step, A.rules);
    LAUNCHES(4);
    LAUNCH_CHECK();
    return 0;
}
int sapio_train_network(const SapioWorld *w, void *ws, size_t ws_bytes, const int32_t *d_orgs, int n, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    if ((rc = check_list(w, d_orgs, n))) return rc;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    k_fill_list<<<(n + 255) / 256, 256, 0, st>>>(A.rules.train_list, A.rules.counters + RC_NTRAIN, d_orgs, n);
    LAUNCHES(1);
    return launch_train(*w, A, st);
}
int sapio_reproduce(const SapioWorld *w, void *ws, size_t ws_bytes, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    return launch_reproduce(*w, A, (cudaStream_t)stream);
}
int sapio_spawn(const SapioWorld *w, void *ws, size_t ws_bytes, const int32_t *d_slots, const double *d_pos, int n, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    if (n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int blocks = (n + REPRO_WARPS - 1) / REPRO_WARPS, cap = (sm_count() > 0 ? sm_count() : 148) * 8;
    if (blocks > cap) blocks = cap;
    k_spawn<<<blocks, REPRO_WARPS * 32, 0, st>>>(*w, d_slots, d_pos, n);
    k_spawn_finish<<<1, 1, 0, st>>>(*w, n);
    LAUNCH_CHECK();
    return 0;
}
int sapio_sensor_view(const SapioWorld *w, void *ws, size_t ws_bytes, int org, int cell, uint32_t *d_out, int cap, int32_t *d_count, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    rc = launch_grid_only(*w, A.step, st); if (rc) return rc;
    k_sensor_view<<<1, 32, 0, st>>>(*w, A.step, org, cell, d_out, cap, d_count);
    LAUNCH_CHECK();
    return 0;
}

int sapio_tick(const SapioWorld *w, void *ws, size_t ws_bytes, uint32_t phases, int n_ticks, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    for (int t = 0; t < n_ticks; t++) {       // inside one call nothing but the tick touches the world: ticks after the first reuse the step's grid
        if ((rc = launch_tick(*w, A, phases, nullptr, (cudaStream_t)stream))) return rc;
        if (phases & SAPIO_PH_STEP) phases |= SAPIO_PH_REUSE_GRID;
    }
    return 0;
}

int sapio_events(const SapioWorld *w, void *ws, size_t ws_bytes, float drought, float poison, float random_cells, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    SapioControl c = {drought, poison, random_cells, 0};
    CUDA_TRY(cudaMemcpyAsync(A.frame.control, &c, sizeof(c), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));                       // `c` lives on this stack frame
    k_events<<<stream_grid((long)w->p.n_org * MAXC), 256, 0, st>>>(*w, A.frame.control);
    k_event_food<<<1, 256, 0, st>>>(*w, A.frame.control);
    LAUNCHES(2);
    LAUNCH_CHECK();
    return 0;
}

int sapio_tick_host(const SapioWorld *w, void *ws, size_t ws_bytes, uint32_t phases, int n_ticks, const SapioControl *h_control,
                    SapioFrameHeader *h_header, float *h_cells, int cell_cap, float *h_dead, int dead_cap, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    if (!h_control || !h_header) return fail(SAPIO_E_ARG, "null host buffer%s");
    cudaStream_t st = (cudaStream_t)stream;
    const long nb = (long)w->p.n_org * MAXC + w->p.n_dead;
    const int ccap = h_cells ? (cell_cap < w->p.n_org * MAXC ? cell_cap : w->p.n_org * MAXC) : 0;
    const int dcap = h_dead ? (dead_cap < w->p.n_dead ? dead_cap : w->p.n_dead) : 0;
    CUDA_TRY(cudaMemcpyAsync(A.frame.control, h_control, sizeof(SapioControl), cudaMemcpyHostToDevice, st));
    for (int t = 0; t < n_ticks; t++) {
        if ((rc = launch_tick(*w, A, phases, A.frame.control, st))) return rc;
        if ((phases & SAPIO_PH_STEP) && !(h_control->random_cells > 0.f)) phases |= SAPIO_PH_REUSE_GRID;      // event food is not in the step's grid
    }
    prof_mark(st, PT_FRAME);
    CUDA_TRY(cudaMemsetAsync(A.frame.counts, 0, 256, st));
    k_pack_frame<<<stream_grid(nb), 256, 0, st>>>(*w, A.frame, ccap, dcap);
    k_frame_header<<<1, 1, 0, st>>>(*w, A.frame);
    LAUNCHES(2);
    CUDA_TRY(cudaMemcpyAsync(h_header, A.frame.header, sizeof(SapioFrameHeader), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    int nc = h_header->n_cells < ccap ? h_header->n_cells : ccap, nd = h_header->n_dead < dcap ? h_header->n_dead : dcap;
    if (nc > 0) CUDA_TRY(cudaMemcpyAsync(h_cells, A.frame.cells, (size_t)nc * sizeof(float4), cudaMemcpyDeviceToHost, st));
    if (nd > 0) CUDA_TRY(cudaMemcpyAsync(h_dead, A.frame.dead, (size_t)nd * sizeof(float4), cudaMemcpyDeviceToHost, st));
    prof_mark(st, PT_END);
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

size_t sapio_organism_record_bytes(const SapioParams *p) {
    if (!p) return 0;
    MigTable T; mig_table(*p, nullptr, T);
    return T.record_bytes;
}
static int mig_check(const SapioWorld *w, const int32_t *d_slots, int n, const void *d_buf, size_t buf_bytes, MigTable &T) {
    int rc = check(w, nullptr, 0, nullptr); if (rc) return rc;
    if ((rc = check_list(w, d_slots, n))) return rc;
    mig_table(w->p, w, T);
    if (T.n > MIG_MAX_FIELDS) return fail(SAPIO_E_ARG, "field table overflow%s");
    if (n && (!d_buf || buf_bytes < (size_t)n * T.record_bytes)) return fail(SAPIO_E_ARG, "record buffer too small%s");
    return 0;
}
int sapio_pack_organisms(const SapioWorld *w, const int32_t *d_slots, int n, void *d_buf, size_t buf_bytes, void *stream) {
    MigTable T; int rc = mig_check(w, d_slots, n, d_buf, buf_bytes, T); if (rc) return rc;
    if (n == 0) return 0;
    k_org_records<true><<<n < 1184 ? n : 1184, 256, 0, (cudaStream_t)stream>>>(T, d_slots, n, (char *)d_buf, w->g_next_uid);
    LAUNCHES(1); LAUNCH_CHECK();
    return 0;
}
int sapio_unpack_organisms(const SapioWorld *w, const int32_t *d_slots, int n, const void *d_buf, size_t buf_bytes, void *stream) {
    MigTable T; int rc = mig_check(w, d_slots, n, d_buf, buf_bytes, T); if (rc) return rc;
    if (n == 0) return 0;
    k_org_records<false><<<n < 1184 ? n : 1184, 256, 0, (cudaStream_t)stream>>>(T, d_slots, n, (char *)const_cast<void *>(d_buf), w->g_next_uid);
    k_uid_advance<<<1, 1, 0, (cudaStream_t)stream>>>(w->g_next_uid, n);
    LAUNCHES(2); LAUNCH_CHECK();
    return 0;
}
int sapio_release_organisms(const SapioWorld *w, const int32_t *d_slots, int n, void *stream) {
    int rc = check(w, nullptr, 0, nullptr); if (rc) return rc;
    if ((rc = check_list(w, d_slots, n))) return rc;
    if (n == 0) return 0;
    k_mig_release<<<n < 4096 ? n : 4096, MAXC, 0, (cudaStream_t)stream>>>(*w, d_slots, nullptr, n);
    k_mig_scrub<<<stream_grid(2 * (long)w->p.x_cap), 256, 0, (cudaStream_t)stream>>>(*w);
    LAUNCHES(2); LAUNCH_CHECK();
    return 0;
}

// ---- spatial slabs (csrc/halo.cuh, sapiogenesis_b200/slabs.py) ----------------------------------------------------------------------
int sapio_rules_part(const SapioWorld *w, void *ws, size_t ws_bytes, uint32_t phases, int part, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    return launch_rules(*w, A, phases, (cudaStream_t)stream, part);
}
int sapio_post_advance(const SapioWorld *w, void *ws, size_t ws_bytes, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    return launch_post(*w, A, (cudaStream_t)stream, 1, 1);
}
int sapio_slab_offsets(const SapioWorld *w, void *ws, size_t ws_bytes, int64_t *out4) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    out4[0] = (char *)(A.rules.gas_scan + (w->p.n_org - 1)) - (char *)ws;      /* this slab's composed gas map (a, b) */
    out4[1] = (char *)A.rules.refund_sum - (char *)ws; out4[2] = (char *)A.rules.disp_prod - (char *)ws; out4[3] = (char *)A.step.pop - (char *)ws;
    return 0;
}
int sapio_halo_pack(const SapioWorld *w, void *ws, size_t ws_bytes, float x_lo, float x_hi, void *d_buf, size_t buf_bytes, void *d_dead, size_t dead_bytes, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    if (!w->p.slab_mode) return fail(SAPIO_E_ARG, "sapio_halo_pack: not a slab world%s");
    cudaStream_t st = (cudaStream_t)stream;
    SapioMigTable T; sapio_mig_table(&w->p, w, &T);
    k_halo_flag<<<stream_grid(w->p.n_org + 1), 256, 0, st>>>(*w, x_lo, x_hi, A.repro.req_flag);          // the reproduction scratch is idle here
    scan_exclusive_i32(A.repro.req_flag, A.repro.req_off, w->p.n_org + 1, A.step.scan_parts, st);
    k_halo_choose<<<stream_grid(w->p.n_org), 256, 0, st>>>(*w, A.mig, A.repro.req_flag, A.repro.req_off, SAPIO_MIG_MAX_RECORDS);
    k_mig_sizes<<<1, 256, 0, st>>>(*w, T, A.mig, (char *)d_buf, buf_bytes);
    k_mig_pack<<<148 * 4, 256, 0, st>>>(*w, T, A.mig, (char *)d_buf);
    k_halo_dead_header<<<1, 1, 0, st>>>((char *)d_dead);
    k_halo_pack_dead<<<stream_grid(w->p.n_dead), 256, 0, st>>>(*w, A.rules, x_lo, x_hi, (char *)d_dead, dead_bytes);
    LAUNCHES(5); LAUNCH_CHECK();
    return 0;
}
int sapio_halo_unpack(const SapioWorld *w, void *ws, size_t ws_bytes, int side, void *d_buf, size_t buf_bytes, void *d_dead, size_t dead_bytes, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    if (!w->p.slab_mode || side < 0 || side > 1) return fail(SAPIO_E_ARG, "sapio_halo_unpack: not a slab world, or side not 0 / 1%s");
    cudaStream_t st = (cudaStream_t)stream;
    SapioMigTable T; sapio_mig_table(&w->p, w, &T);
    k_halo_assign_kept<<<16, 256, 0, st>>>(*w, T, A.mig, A.halo, side, (const char *)d_buf);
    k_halo_assign<<<1, 32, 0, st>>>(*w, T, A.mig, A.halo, side, (const char *)d_buf);
    k_halo_unpack<<<148 * 4, 256, 0, st>>>(*w, T, A.mig, (char *)d_buf);
    k_halo_assign_dead<<<1, 32, 0, st>>>(*w, A.rules, A.halo, side, (const char *)d_dead);
    k_halo_unpack_dead<<<stream_grid(w->p.n_dead), 256, 0, st>>>(*w, A.rules, A.halo, (const char *)d_dead);
    LAUNCHES(4); LAUNCH_CHECK();
    (void)buf_bytes; (void)dead_bytes;
    return 0;
}
int sapio_halo_sweep(const SapioWorld *w, void *ws, size_t ws_bytes, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    if (!w->p.slab_mode) return fail(SAPIO_E_ARG, "sapio_halo_sweep: not a slab world%s");
    cudaStream_t st = (cudaStream_t)stream;
    k_halo_sweep<<<stream_grid(w->p.n_org > w->p.n_dead ? w->p.n_org : w->p.n_dead), 256, 0, st>>>(*w, A.rules, A.halo);
    k_mig_scrub<<<stream_grid(2 * (long)w->p.x_cap), 256, 0, st>>>(*w);
    LAUNCHES(2); LAUNCH_CHECK();
    return 0;
}
int sapio_slab_gas_begin(const SapioWorld *w, void *ws, size_t ws_bytes, const double *d_maps, int n, int rank, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    if (rank < 0 || rank >= n) return fail(SAPIO_E_ARG, "sapio_slab_gas_begin: rank out of range%s");
    k_slab_gas_begin<<<1, 1, 0, (cudaStream_t)stream>>>(*w, A.rules, d_maps, rank);
    LAUNCHES(1); LAUNCH_CHECK();
    return 0;
}
int sapio_slab_gas_finish(const SapioWorld *w, const double *d_maps, int n, const double *d_refunds, void *stream) {
    int rc = check(w, nullptr, 0, nullptr); if (rc) return rc;
    k_slab_gas_finish<<<1, 1, 0, (cudaStream_t)stream>>>(*w, d_maps, n, d_refunds);
    LAUNCHES(1); LAUNCH_CHECK();
    return 0;
}
int sapio_slab_disperse_finish(const SapioWorld *w, const double *d_prods, int n, void *stream) {
    int rc = check(w, nullptr, 0, nullptr); if (rc) return rc;
    k_slab_disperse_finish<<<1, 1, 0, (cudaStream_t)stream>>>(*w, d_prods, n);
    LAUNCHES(1); LAUNCH_CHECK();
    return 0;
}
int sapio_slab_population(const SapioWorld *w, const int32_t *d_counts, int n, void *stream) {
    int rc = check(w, nullptr, 0, nullptr); if (rc) return rc;
    k_slab_population<<<1, 1, 0, (cudaStream_t)stream>>>(*w, d_counts, n);
    LAUNCHES(1); LAUNCH_CHECK();
    return 0;
}

// ---- island migration on the device (csrc/migrate2.cuh, include/sapio_migrate.h) ---------------------------------------------
size_t sapio_migrate_buffer_bytes(const SapioParams *p, int count, int bytes_per_organism) {
    if (!p || count < 0) return 0;
    if (count > SAPIO_MIG_MAX_RECORDS) count = SAPIO_MIG_MAX_RECORDS;
    return ((sizeof(SapioMigBufHeader) + 15) & ~size_t(15)) + (size_t)count * (size_t)(bytes_per_organism > 0 ? bytes_per_organism : 65536);
}
int sapio_migrate_out(const SapioWorld *w, void *ws, size_t ws_bytes, int count, int rank, void *d_buf, size_t buf_bytes, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    if (!d_buf || buf_bytes < sizeof(SapioMigBufHeader) + 64) return fail(SAPIO_E_ARG, "sapio_migrate_out: buffer too small%s");
    if (count < 0 || count > SAPIO_MIG_MAX_RECORDS) return fail(SAPIO_E_ARG, "sapio_migrate_out: count out of range%s");
    cudaStream_t st = (cudaStream_t)stream;
    SapioMigTable T; sapio_mig_table(&w->p, w, &T);
    k_mig_choose<<<1, 256, 0, st>>>(*w, A.mig, count, rank);
    k_mig_sizes<<<1, 256, 0, st>>>(*w, T, A.mig, (char *)d_buf, buf_bytes);
    k_mig_pack<<<count > 0 ? count : 1, 256, 0, st>>>(*w, T, A.mig, (char *)d_buf);
    k_mig_release<<<count > 0 ? count : 1, MAXC, 0, st>>>(*w, A.mig.emig, A.mig.counters + MC_PACKED, 0);
    k_mig_scrub<<<stream_grid(2 * (long)w->p.x_cap), 256, 0, st>>>(*w);
    LAUNCHES(5); LAUNCH_CHECK();
    return 0;
}
int sapio_migrate_in(const SapioWorld *w, void *ws, size_t ws_bytes, void *d_buf, size_t buf_bytes, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    if (!d_buf || buf_bytes < sizeof(SapioMigBufHeader)) return fail(SAPIO_E_ARG, "sapio_migrate_in: buffer too small%s");
    cudaStream_t st = (cudaStream_t)stream;
    SapioMigTable T; sapio_mig_table(&w->p, w, &T);
    const int cap = SAPIO_MIG_MAX_RECORDS;                              // the count lives in the buffer, on the device
    k_org_rects<<<stream_grid(w->p.n_org), 256, 0, st>>>(*w, nullptr);
    k_mig_prepare<<<(cap + 127) / 128, 128, 0, st>>>(*w, T, A.mig, (const char *)d_buf);
    k_mig_viable<<<SAPIO_MIG_MAX_RECORDS < 592 ? SAPIO_MIG_MAX_RECORDS : 592, MIG_VTHREADS, 0, st>>>(*w, A.mig, (const char *)d_buf);
    k_mig_resolve<<<1, 32, 0, st>>>(*w, A.mig, (const char *)d_buf);
    k_mig_unpack<<<148, 256, 0, st>>>(*w, T, A.mig, (char *)d_buf);
    k_mig_clear_dead<<<stream_grid(w->p.n_dead), 256, 0, st>>>(*w, A.mig);
    k_mig_finish<<<1, 1, 0, st>>>(*w, A.mig);
    LAUNCHES(7); LAUNCH_CHECK();
    return 0;
}
int sapio_migrate_counts(const SapioWorld *w, void *ws, size_t ws_bytes, int32_t *h_out8, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(h_out8, A.mig.counters, 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

// ---- pipelined host exchange: the download of frame k runs on a copy stream while tick k + 1 computes --------------------
struct HostFrame {
    cudaEvent_t packed = nullptr, copied = nullptr;
    int sent_cells = 0, sent_dead = 0;            // what the queued download covers
    SapioFrameHeader *h_header = nullptr; float *h_cells = nullptr, *h_dead = nullptr; int ccap = 0, dcap = 0;
    const float4 *d_cells = nullptr, *d_dead = nullptr;
};
struct HostPipe {
    bool ready = false;
    cudaStream_t copy = nullptr;
    HostFrame fr[2];
    int head = 0, count = 0;                      // ring of frames in flight (oldest first)
    long seq = 0;                                 // frames begun: parity picks the device frame buffer
    int last_cells = -1, last_dead = -1;          // counts of the last collected frame (host side)
};
static HostPipe g_pipe[64];

int sapio_tick_host_begin(const SapioWorld *w, void *ws, size_t ws_bytes, uint32_t phases, int n_ticks, const SapioControl *h_control,
                          SapioFrameHeader *h_header, float *h_cells, int cell_cap, float *h_dead, int dead_cap, void *stream) {
    AllWs A; int rc = check(w, ws, ws_bytes, &A); if (rc) return rc;
    if (!h_control || !h_header) return fail(SAPIO_E_ARG, "null host buffer%s");
    int dev = 0; CUDA_TRY(cudaGetDevice(&dev));
    HostPipe &P = g_pipe[dev & 63];
    if (P.count == 2) return fail(SAPIO_E_ARG, "sapio_tick_host_begin: two frames are already in flight (collect one with sapio_tick_host_end)%s");
    if (!P.ready) {
        CUDA_TRY(cudaStreamCreateWithFlags(&P.copy, cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++) {
            CUDA_TRY(cudaEventCreateWithFlags(&P.fr[k].packed, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&P.fr[k].copied, cudaEventDisableTiming));
        }
        P.ready = true;
    }
    HostFrame &H = P.fr[(P.head + P.count) & 1];
    FrameWs &F = (P.seq & 1) ? A.frame2 : A.frame;                   // the frame before last used this buffer and has been collected
    cudaStream_t st = (cudaStream_t)stream;
    const long nb = (long)w->p.n_org * MAXC + w->p.n_dead;
    const int ccap = h_cells ? (cell_cap < w->p.n_org * MAXC ? cell_cap : w->p.n_org * MAXC) : 0;
    const int dcap = h_dead ? (dead_cap < w->p.n_dead ? dead_cap : w->p.n_dead) : 0;
    CUDA_TRY(cudaMemcpyAsync(A.frame.control, h_control, sizeof(SapioControl), cudaMemcpyHostToDevice, st));
    for (int t = 0; t < n_ticks; t++) {
        if ((rc = launch_tick(*w, A, phases, A.frame.control, st))) return rc;
        if ((phases & SAPIO_PH_STEP) && !(h_control->random_cells > 0.f)) phases |= SAPIO_PH_REUSE_GRID;      // event food is not in the step's grid
    }
    prof_mark(st, PT_FRAME);
    CUDA_TRY(cudaMemsetAsync(F.counts, 0, 256, st));
    k_pack_frame<<<stream_grid(nb), 256, 0, st>>>(*w, F, ccap, dcap);
    k_frame_header<<<1, 1, 0, st>>>(*w, F);
    LAUNCHES(2);
    prof_mark(st, PT_END);
    CUDA_TRY(cudaEventRecord(H.packed, st));
    CUDA_TRY(cudaStreamWaitEvent(P.copy, H.packed, 0));
    // the counts are only known on the device: download what the last collected frame held plus a margin; _end fetches a rest
    auto guess = [](int last, int cap) { if (last < 0) return cap; long g = (long)last + last / 16 + 8192; return (int)(g < cap ? g : cap); };
    H.sent_cells = guess(P.last_cells, ccap); H.sent_dead = guess(P.last_dead, dcap);
    CUDA_TRY(cudaMemcpyAsync(h_header, F.header, sizeof(SapioFrameHeader), cudaMemcpyDeviceToHost, P.copy));
    if (H.sent_cells > 0) CUDA_TRY(cudaMemcpyAsync(h_cells, F.cells, (size_t)H.sent_cells * sizeof(float4), cudaMemcpyDeviceToHost, P.copy));
    if (H.sent_dead > 0) CUDA_TRY(cudaMemcpyAsync(h_dead, F.dead, (size_t)H.sent_dead * sizeof(float4), cudaMemcpyDeviceToHost, P.copy));
    CUDA_TRY(cudaEventRecord(H.copied, P.copy));
    H.h_header = h_header; H.h_cells = h_cells; H.h_dead = h_dead; H.ccap = ccap; H.dcap = dcap; H.d_cells = F.cells; H.d_dead = F.dead;
    P.count++; P.seq++;
    return 0;
}
int sapio_tick_host_end(void) {
    int dev = 0; CUDA_TRY(cudaGetDevice(&dev));
    HostPipe &P = g_pipe[dev & 63];
    if (P.count == 0) return fail(SAPIO_E_ARG, "sapio_tick_host_end without sapio_tick_host_begin%s");
    HostFrame &H = P.fr[P.head];
    CUDA_TRY(cudaEventSynchronize(H.copied));
    const int nc = H.h_header->n_cells < H.ccap ? H.h_header->n_cells : H.ccap, nd = H.h_header->n_dead < H.dcap ? H.h_header->n_dead : H.dcap;
    bool rest = false;
    if (nc > H.sent_cells) { CUDA_TRY(cudaMemcpyAsync(H.h_cells + 4 * (size_t)H.sent_cells, H.d_cells + H.sent_cells, (size_t)(nc - H.sent_cells) * sizeof(float4), cudaMemcpyDeviceToHost, P.copy)); rest = true; }
    if (nd > H.sent_dead) { CUDA_TRY(cudaMemcpyAsync(H.h_dead + 4 * (size_t)H.sent_dead, H.d_dead + H.sent_dead, (size_t)(nd - H.sent_dead) * sizeof(float4), cudaMemcpyDeviceToHost, P.copy)); rest = true; }
    if (rest) CUDA_TRY(cudaStreamSynchronize(P.copy));
    P.last_cells = nc; P.last_dead = nd;
    P.head ^= 1; P.count--;
    return 0;
}

int sapio_profile(int enable) { if (!enable && g_prof.on) prof_collect(); g_prof.on = enable != 0; return 0; }
int sapio_profile_stages(void) { return PT_COUNT - 1; }
const char *sapio_profile_name(int i) { return (i >= 0 && i < PT_COUNT) ? PROF_NAMES[i] : ""; }
int sapio_profile_read(double *ms_out, int64_t *count_out, int reset) {
    prof_collect();
    for (int i = 0; i < PT_COUNT - 1; i++) { if (ms_out) ms_out[i] = g_prof.ms[i]; if (count_out) count_out[i] = g_prof.count[i]; }
    if (reset) for (int i = 0; i < PT_COUNT; i++) { g_prof.ms[i] = 0; g_prof.count[i] = 0; }
    return 0;
}
uint64_t sapio_launch_count(void) { return g_prof.launches; }
int sapio_solver_info(int cls, int32_t *out6) {
    if (cls < 0 || cls >= ORG_CLASSES || !out6) return fail(SAPIO_E_ARG, "sapio_solver_info: class out of range%s");
    int info[4]; ORG_INFO[cls](info);
    out6[0] = info[0]; out6[1] = info[1]; out6[2] = info[2]; out6[3] = info[3];
    out6[4] = g_plan.ready ? g_plan.org_blocks[cls] : 0; out6[5] = g_plan.ready ? g_plan.sms : 0;
    return 0;
}
int sapio_debug_flags(int f) { const int old = g_debug_flags; g_debug_flags = f; return old; }      // honoured by builds with -DSAPIO_DEBUG_FLAGS only
int sapio_set_coupled_pack_from(int n) { const int old = g_force_fallback ? 2147483647 : 2400; g_force_fallback = n == 2147483647; return old; }

}  // extern "C"
