// coupled.cuh -- organisms and dead cells that share a cross contact: cooperative kernels (own translation unit: coupled.cu)
#pragma once
#include <cooperative_groups.h>
#include "orgsolve.cuh"
namespace cg = cooperative_groups;
// ------------------------------------------------------------------------------------------------------------------
// coupled path: cooperative kernel
// ------------------------------------------------------------------------------------------------------------------
// dependency levels of the cross contacts: level[c] = 1 + max(level of the previous contact on either body), relaxed upwards
// from all ones to the fixed point (== longest chain ending in c; the relaxation is monotone, so any sweep order ends there).
// A short list is done by one CTA between block barriers while the rest of the grid waits once; a long one by the whole grid
// with one grid barrier per sweep (three rotating "changed" flags: sweep r clears the flag of sweep r + 1 and sets its own).
#define LEVELS_ONE_CTA 4096
__device__ __forceinline__ bool cross_levels_relax(const Workspace &W, int c) {
    const int pa = W.nx_prev_a[c], pb = W.nx_prev_b[c];
    const int L = 1 + max(pa >= 0 ? W.nx_level[pa] : 0, pb >= 0 ? W.nx_level[pb] : 0);
    if (L <= W.nx_level[c]) return false;
    W.nx_level[c] = L; atomicMax(&W.flags[FL_MAXLEVEL], L);
    return true;
}
__device__ __forceinline__ void cross_levels(cg::grid_group &grid, const Workspace &W, int nx, int gtid, int gthreads) {
    if (nx <= LEVELS_ONE_CTA) {
        if (blockIdx.x == 0) {
            __shared__ int s_changed;
            for (;;) {
                __syncthreads();
                if (threadIdx.x == 0) s_changed = 0;
                __syncthreads();
                bool ch = false;
                for (int c = threadIdx.x; c < nx; c += blockDim.x) ch |= cross_levels_relax(W, c);
                if (ch) s_changed = 1;
                __syncthreads();
                if (!s_changed) break;
            }
        }
        grid.sync();
        return;
    }
    if (gtid == 0) W.flags[FL_RARE] |= SAPIO_RARE_GRID_LEVELS;
    for (int r = 0;; r++) {
        if (gtid == 0) W.flags[FL_CHG0 + (r + 1) % 3] = 0;
        bool ch = false;
        for (int c = gtid; c < nx; c += gthreads) ch |= cross_levels_relax(W, c);
        if (ch) W.flags[FL_CHG0 + r % 3] = 1;
        grid.sync();
        if (!W.flags[FL_CHG0 + r % 3]) break;
    }
}

template <bool WARM>
__device__ __forceinline__ void cross_level_pass(const SapioWorld &w, const Workspace &W, int L, int nx, float dt_coef, int gtid, int gthreads) {
    const int NC = w.p.n_org * MAXC, ND = w.p.n_dead;
    for (int c = gtid; c < nx; c += gthreads) {
        if (W.nx_level[c] != L) continue;
        uint64_t key = W.nx_key[c];
        uint32_t a = (uint32_t)(key >> 32), b = (uint32_t)key;
        GRef A, B; gload(w, a, NC, ND, A); gload(w, b, NC, ND, B);
        float2 n = W.nx_n[c];
        if (WARM) contact_warm(A.b, B.b, n.x, n.y, W.nx_jn[c], W.nx_jt[c], dt_coef);
        else {
            float jn = W.nx_jn[c], jt = W.nx_jt[c], jb = W.nx_jbias[c];
            contact_apply(A.b, B.b, n.x, n.y, W.nx_bias[c], W.nx_bounce[c], A.u * B.u, jn, jt, jb);
            W.nx_jn[c] = jn; W.nx_jt[c] = jt; W.nx_jbias[c] = jb;
        }
        gstore(w, a, NC, A); gstore(w, b, NC, B);
    }
}

// Warm start of the cross contacts (cpArbiterApplyCachedImpulse), exact: cached impulses do not depend on the velocities, so every
// body receives the sum of its contacts' impulses at once, summed in double (OrgCtx::warm_start has the reasons). Dead cells are
// done here, one thread per dead body with partners; living cells get theirs inside their organism's warm start (CrossView).
__device__ __forceinline__ CrossView cross_view(const Workspace &W) {
    return CrossView{W.off_all, W.adj_c, W.nx_key, W.nx_n, W.nx_jn, W.nx_jt, W.flags[FL_NX]};
}
__device__ __forceinline__ void cross_warm_dead(const SapioWorld &w, const Workspace &W, float dt_coef, int tid, int nthreads) {
    const int NC = w.p.n_org * MAXC, ND = w.p.n_dead, nx = W.flags[FL_NX];
    const int nh = min(W.flags[FL_NHIT], 2 * w.p.x_cap);
    const double dt = (double)dt_coef;
    for (int h = tid; h < nh; h += nthreads) {
        const uint32_t body = (uint32_t)W.hit_list[h].x;
        if ((int)body < NC || (int)body >= NC + ND) continue;
        const int d = (int)body - NC;
        double ax = 0.0, ay = 0.0, at = 0.0;
        for (int q = W.off_all[body], e = W.off_all[body + 1]; q < e; q++) {
            const int c = W.adj_c[q];
            if (c < 0 || c >= nx) continue;
            const float2 nn = W.nx_n[c];
            const double jn = (double)W.nx_jn[c] * dt, jt = (double)W.nx_jt[c] * dt, nxx = (double)nn.x, nyy = (double)nn.y;
            const double jx = nxx * jn - nyy * jt, jy = nxx * jt + nyy * jn;
            if ((uint32_t)(W.nx_key[c] >> 32) == body) { ax -= jx; ay -= jy; } else { ax += jx; ay += jy; }
            at += jt;
        }
        const float r = w.d_size[d] * 0.5f, minv = 1.0f / w.d_mass[d], iinv = 1.0f / (0.5f * w.d_mass0[d] * r * r);
        float2 v = reinterpret_cast<const float2 *>(w.d_vel)[d];
        reinterpret_cast<float2 *>(w.d_vel)[d] = make_float2((float)fma(ax, (double)minv, (double)v.x), (float)fma(ay, (double)minv, (double)v.y));
        w.d_angvel[d] = (float)((double)w.d_angvel[d] - (double)(iinv * r) * at);
    }
}

// Coupled organisms go through the same lane-packed, size-classed solver as the others, one pass at a time between the grid
// barriers. A warp owns a slab of shared memory that holds one warp-item of any class up to 48 rows (8/4/2/2/1/1 organisms);
// the two largest classes take two adjacent slabs (even warps only, in their own sub-phase).
#define COUPLED_WARPS 8


template <int CM> struct CpBytes { static constexpr size_t v = OrgCfg<CM>::NG * sizeof(OrgS<CM>); };
constexpr size_t cp_max(size_t a, size_t b) { return a > b ? a : b; }
constexpr size_t CP_SLAB = (cp_max(cp_max(cp_max(CpBytes<8>::v, CpBytes<16>::v), cp_max(CpBytes<24>::v, CpBytes<32>::v)), cp_max(CpBytes<40>::v, CpBytes<48>::v)) + 15) & ~size_t(15);
static_assert(2 * CP_SLAB >= sizeof(OrgS<64>) && 2 * CP_SLAB >= sizeof(OrgS<56>), "two slabs must hold the largest organism");
static_assert(sizeof(OrgS<MAXC>) <= COUPLED_STATE_BYTES, "COUPLED_STATE_BYTES too small");
enum { CP_STAGE_A = 0, CP_WARM = 1, CP_ITER = 2, CP_LAST = 3 };

template <int CM>
__device__ void coupled_item(const SapioWorld &w, const Workspace &W, const StepK &K, unsigned char *slab, int item, int kind, float dt_coef) {
    constexpr int G = OrgCfg<CM>::G, NG = OrgCfg<CM>::NG, cls = (CM - 1) >> 3;
    typedef OrgCtx<CM, G> Ctx;
    const int lane = lane_id();
    OrgS<CM> &s = reinterpret_cast<OrgS<CM> *>(slab)[lane / G];
    const int start = W.flags[FL_CCLS0 + cls], end = W.flags[FL_CCLS0 + cls + 1];
    const int idx = start + item * NG + lane / G;                                 // position in the class-sorted coupled list
    const int o = idx < end ? W.coupled_list[idx] : -1;
    Ctx c(s, w, K, o);
    OrgS<CM> *g = (o >= 0 && idx < W.cp_cap) ? reinterpret_cast<OrgS<CM> *>(W.cp_state + (size_t)idx * COUPLED_STATE_BYTES) : nullptr;
    if (kind == CP_STAGE_A) {   // detect the intra contacts and take their bounce before any velocity changes; park the pre-stepped state
        c.load(); c.prestep_pairs(); c.match_prev(); c.schedule(); c.template prestep_contacts<true>();
        if (dt_coef == 0.f) c.zero_pin_acc();                                     // no warm start will consume the cached pin impulses
        if (o >= 0) {
            c.store_contacts(W.ic_bounce, W.ic_jbias);
            if (g) c.save_static(g);
            if (c.sl == 0) atomicAdd(&W.flags[FL_NIC], s.nic);
        }
        __syncwarp();
        return;
    }
    const bool use_park = __all_sync(FULL, o < 0 || g != nullptr);                // warp-uniform: the two paths are collective
    c.load();
    if (use_park) c.restore_static(g);
    else { c.prestep_pairs(); c.schedule(); c.template prestep_contacts<false>(); if (kind != CP_WARM) c.zero_pin_acc(); }   // the cached total stays in global memory
    c.load_contact_state(W.ic_bounce, W.ic_jbias);
    const int nsteps_max = Ctx::wmax(s.nsteps), tmax = 2 * Ctx::wmax(c.n) - 3;
    if (kind == CP_WARM) { const CrossView X = cross_view(W); c.warm_start(dt_coef, &X); }
    else { c.template contacts_pass<false>(0.f, nsteps_max); c.template joints_pass<false>(0.f, tmax); }
    if (o >= 0) {
        c.store_vel(!use_park || kind == CP_LAST);
        if (use_park) c.save_pairs(g);
        if (kind != CP_WARM) c.store_contacts(W.ic_bounce, W.ic_jbias);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(COUPLED_WARPS * 32) k_coupled(WORLD_ARG, Workspace W, StepK K) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nx = W.flags[FL_NX];
    if (nx == 0) return;                                       // uniform over the grid: no barrier has been entered
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    const int warp = threadIdx.x >> 5;
    unsigned char *slab = smem_raw + (size_t)warp * CP_SLAB;
    const float dt_coef = w.g_stepped[0] ? 1.0f : 0.0f;
    if (gtid == 0) W.flags[FL_RARE] |= SAPIO_RARE_PACKED_COUPLED;
    cross_levels(grid, W, nx, gtid, gthreads);
    const int maxL = max(W.flags[FL_MAXLEVEL], 1);
    // warp-items per class (prefix): classes 0..5 by every warp, classes 6..7 by the even warps on a double slab
    int ibase[ORG_CLASSES + 1];
    ibase[0] = 0;
#pragma unroll
    for (int c = 0; c < ORG_CLASSES; c++) {
        const int cnt = W.flags[FL_CCLS0 + c + 1] - W.flags[FL_CCLS0 + c], ng = c == 0 ? 8 : c == 1 ? 4 : c <= 3 ? 2 : 1;
        ibase[c + 1] = ibase[c] + (cnt + ng - 1) / ng;
    }
    // Work is pulled, heaviest class first, from a pair of counters: a phase ends at a grid barrier, so what counts is when its last
    // warp finishes. Two pairs alternate; a phase zeroes the pair of the next one (idle since the barrier before this phase; the
    // first pair starts at zero with the flags block).
    int phase_no = 0;
    auto pull = [&](int counter) { int t = 0; if (lane_id() == 0) t = atomicAdd(&W.flags[counter], 1); return __shfl_sync(FULL, t, 0); };
    auto org_phase = [&](int kind) {
        const int ca = FL_CWORK0 + 2 * (phase_no & 1), cb = ca + 1;
        if (gtid == 0) { W.flags[FL_CWORK0 + 2 * ((phase_no + 1) & 1)] = 0; W.flags[FL_CWORK0 + 2 * ((phase_no + 1) & 1) + 1] = 0; }
        phase_no++;
        for (;;) {
            const int it = ibase[6] - 1 - pull(ca);
            if (it < 0) break;
            if (it < ibase[1]) coupled_item<8>(w, W, K, slab, it - ibase[0], kind, dt_coef);
            else if (it < ibase[2]) coupled_item<16>(w, W, K, slab, it - ibase[1], kind, dt_coef);
            else if (it < ibase[3]) coupled_item<24>(w, W, K, slab, it - ibase[2], kind, dt_coef);
            else if (it < ibase[4]) coupled_item<32>(w, W, K, slab, it - ibase[3], kind, dt_coef);
            else if (it < ibase[5]) coupled_item<40>(w, W, K, slab, it - ibase[4], kind, dt_coef);
            else coupled_item<48>(w, W, K, slab, it - ibase[5], kind, dt_coef);
        }
        if (ibase[8] > ibase[6]) {                             // uniform over the grid
            __syncthreads();                                   // the odd warps' slabs are borrowed now
            if (!(warp & 1))
                for (;;) {
                    const int it = ibase[8] - 1 - pull(cb);
                    if (it < ibase[6]) break;
                    if (it < ibase[7]) coupled_item<56>(w, W, K, slab, it - ibase[6], kind, dt_coef);
                    else coupled_item<64>(w, W, K, slab, it - ibase[7], kind, dt_coef);
                }
            __syncthreads();
        }
    };
    org_phase(CP_STAGE_A);
    grid.sync();
    // cached impulses: cross contacts by level, then each coupled organism's own arbiters and constraints
    if (dt_coef != 0.f) {
        cross_warm_dead(w, W, dt_coef, gtid, gthreads);       // disjoint from what the organisms touch: no barrier in between
        org_phase(CP_WARM);
        grid.sync();
    }
    const int iterations = w.p.iterations;
    for (int it = 0; it < iterations; it++) {
        for (int L = 1; L <= maxL; L++) { cross_level_pass<false>(w, W, L, nx, 0.f, gtid, gthreads); grid.sync(); }
        org_phase(it == iterations - 1 ? CP_LAST : CP_ITER);
        grid.sync();
    }
}

// The light-load form of the same pass (a few thousand coupled organisms at most): one warp per organism on a full-size tile, four
// warps per SM, every phase on the whole grid. Its phases are short, which is what counts when the barriers outnumber the work;
// k_coupled above wins once the coupled organisms queue up behind the warps (launch_space_step picks by last tick's count).
#define SIMPLE_WARPS 4
__global__ void __launch_bounds__(SIMPLE_WARPS * 32) k_coupled_simple(WORLD_ARG, Workspace W, StepK K) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OrgS<MAXC> &s = reinterpret_cast<OrgS<MAXC> *>(smem_raw)[threadIdx.x >> 5];
    typedef OrgCtx<MAXC, 32> Ctx;
    const int nx = W.flags[FL_NX];
    if (nx == 0) return;                                       // uniform over the grid: no barrier has been entered
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    const int gwarp = gtid >> 5, gwarps = gthreads >> 5;
    const int ncoupled = W.flags[FL_NCOUPLED];
    const float dt_coef = w.g_stepped[0] ? 1.0f : 0.0f;
    cross_levels(grid, W, nx, gtid, gthreads);
    const int maxL = max(W.flags[FL_MAXLEVEL], 1);
    // stage A: coupled organisms detect their intra contacts and take their bounce before any velocity changes
    static_assert(sizeof(OrgS<MAXC>) <= COUPLED_STATE_BYTES, "COUPLED_STATE_BYTES too small");
    auto parked = [&](int q) { return q < W.cp_cap ? reinterpret_cast<OrgS<MAXC> *>(W.cp_state + (size_t)q * COUPLED_STATE_BYTES) : nullptr; };
    for (int q = gwarp; q < ncoupled; q += gwarps) {
        Ctx c(s, w, K, W.coupled_list[q]);
        c.load(); c.prestep_pairs(); c.match_prev(); c.schedule(); c.template prestep_contacts<true>();
        if (dt_coef == 0.f) c.zero_pin_acc();                  // no warm start will consume the cached pin impulses
        c.store_contacts(W.ic_bounce, W.ic_jbias);
        if (OrgS<MAXC> *g = parked(q)) c.save_static(g);
        if (lane_id() == 0) atomicAdd(&W.flags[FL_NIC], s.nic);
        __syncwarp();
    }
    grid.sync();
    // one pass of a coupled organism's own arbiters and constraints between two rounds of cross contacts
    auto org_pass = [&](int q, bool warm, bool last) {
        Ctx c(s, w, K, W.coupled_list[q]);
        c.load();
        OrgS<MAXC> *g = parked(q);
        if (g) c.restore_static(g);
        else { c.prestep_pairs(); c.schedule(); c.template prestep_contacts<false>(); if (!warm) c.zero_pin_acc(); }
        c.load_contact_state(W.ic_bounce, W.ic_jbias);
        if (warm) { const CrossView X = cross_view(W); c.warm_start(dt_coef, &X); }
        else { c.template contacts_pass<false>(0.f, s.nsteps); c.template joints_pass<false>(0.f, 2 * c.n - 3); }
        c.store_vel(!g || last);
        if (g) c.save_pairs(g);
        if (!warm) c.store_contacts(W.ic_bounce, W.ic_jbias);
        __syncwarp();
    };
    // cached impulses: cross contacts by level, then each coupled organism's own arbiters and constraints
    if (dt_coef != 0.f) {
        cross_warm_dead(w, W, dt_coef, gtid, gthreads);       // disjoint from what the organisms touch: no barrier in between
        for (int q = gwarp; q < ncoupled; q += gwarps) org_pass(q, true, false);
        grid.sync();
    }
    const int iterations = w.p.iterations;
    for (int it = 0; it < iterations; it++) {
        for (int L = 1; L <= maxL; L++) { cross_level_pass<false>(w, W, L, nx, 0.f, gtid, gthreads); grid.sync(); }
        for (int q = gwarp; q < ncoupled; q += gwarps) org_pass(q, false, it == iterations - 1);
        grid.sync();
    }
}

