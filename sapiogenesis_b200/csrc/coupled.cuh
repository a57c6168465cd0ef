// coupled.cuh -- organisms and dead cells that share a cross contact: the cooperative fallback kernel (own translation unit: coupled.cu)
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
    W.nx_level[c] = L; atomicMax(&W.flags[FL_MAXLEVEL], L); atomicMax(&W.flags[FL_BIGLEVEL], L);
    return true;
}
template <class Sync>
__device__ __forceinline__ void cross_levels(Sync &sync, const Workspace &W, int nx, int gtid, int gthreads) {
    if (nx <= LEVELS_ONE_CTA) {
        if (blockIdx.x == 0) {
            __shared__ int s_changed;
            for (;;) {
                __syncthreads();
                if (threadIdx.x == 0) s_changed = 0;
                __syncthreads();
                bool ch = false;
                for (int c = threadIdx.x; c < nx; c += blockDim.x) if (W.cn_tier[W.comp_of[c]] < 0) ch |= cross_levels_relax(W, c);
                if (ch) s_changed = 1;
                __syncthreads();
                if (!s_changed) break;
            }
        }
        sync();
        return;
    }
    if (gtid == 0) W.flags[FL_RARE] |= SAPIO_RARE_GRID_LEVELS;
    for (int r = 0;; r++) {
        if (gtid == 0) W.flags[FL_CHG0 + (r + 1) % 3] = 0;
        bool ch = false;
        for (int c = gtid; c < nx; c += gthreads) if (W.cn_tier[W.comp_of[c]] < 0) ch |= cross_levels_relax(W, c);
        if (ch) W.flags[FL_CHG0 + r % 3] = 1;
        sync();
        if (!W.flags[FL_CHG0 + r % 3]) break;
    }
}

template <bool WARM>
__device__ __forceinline__ void cross_level_pass(const SapioWorld &w, const Workspace &W, int L, int nx, float dt_coef, int gtid, int gthreads) {
    const int NC = w.p.n_org * MAXC, ND = w.p.n_dead;
    for (int c = gtid; c < nx; c += gthreads) {
        if (W.nx_level[c] != L || W.cn_tier[W.comp_of[c]] >= 0) continue;         // only the components that no CTA tier holds
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
__device__ __forceinline__ void cross_warm_dead(const SapioWorld &w, const Workspace &W, float dt_coef, int tid, int nthreads) {
    const int NC = w.p.n_org * MAXC, ND = w.p.n_dead, nx = W.flags[FL_NX];
    const int nh = min(W.flags[FL_NHIT], 2 * w.p.x_cap);
    const double dt = (double)dt_coef;
    for (int h = tid; h < nh; h += nthreads) {
        const uint32_t body = (uint32_t)W.hit_list[h].x;
        if ((int)body < NC || (int)body >= NC + ND) continue;
        const int d = (int)body - NC;
        if (!comp_is_big(W, w.p.n_org + d)) continue;
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

// The fallback form of the coupled pass (components.cuh solves every ordinary tick): one warp per organism on a full-size tile, four
// warps per SM, every phase on the whole grid between grid barriers, cross contacts through global memory by dependency level.
// It runs when a connected component exceeds the largest CTA tier (or when a test forces it), and is the second, independent
// implementation the component solver is compared with (tests: fixture `coupled_kernel`).
#define SIMPLE_WARPS 4
#define CLUSTER_CTAS 8                                        // portable cluster size: 32 warps, one organism each
#define CLUSTER_ORGS (CLUSTER_CTAS * SIMPLE_WARPS)
// The body of both launch forms. `sync` is the barrier between phases: a grid barrier of the cooperative launch, or -- when the oversize
// components hold no more organisms than one thread-block cluster has warps, which is what an aged world produces (one cluster of
// 5-20 organisms now and then) -- the hardware barrier of a cluster of 8 CTAs, an order of magnitude cheaper (~170 barriers per tick).
template <class Sync>
__device__ __forceinline__ void coupled_body(const SapioWorld &w, const Workspace &W, const StepK &K, Sync &sync) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OrgS<MAXC> &s = reinterpret_cast<OrgS<MAXC> *>(smem_raw)[threadIdx.x >> 5];
    typedef OrgCtx<MAXC, 32> Ctx;
    const int nx = W.flags[FL_NX];
    if (blockIdx.x == 0 && threadIdx.x == 0) W.flags[FL_RARE] |= SAPIO_RARE_PACKED_COUPLED;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    const int gwarp = gtid >> 5, gwarps = gthreads >> 5;
    const int ncoupled = W.flags[FL_NCOUPLED];
    const float dt_coef = w.g_stepped[0] ? 1.0f : 0.0f;
    // the organisms of the oversize components as a list (order immaterial: within a phase they touch nothing of each other)
    for (int q = gwarp; q < ncoupled; q += gwarps) {
        const int o = W.coupled_list[q];
        if (comp_is_big(W, o) && lane_id() == 0) W.big_list[atomicAdd(&W.flags[FL_BIG_N], 1)] = o;
    }
    cross_levels(sync, W, nx, gtid, gthreads);              // ends with a barrier: the list is complete
    const int nbig = W.flags[FL_BIG_N];
    const int maxL = max(W.flags[FL_BIGLEVEL], 1);
    // stage A: coupled organisms detect their intra contacts and take their bounce before any velocity changes
    static_assert(sizeof(OrgS<MAXC>) <= COUPLED_STATE_BYTES, "COUPLED_STATE_BYTES too small");
    auto parked = [&](int q) { return q < W.cp_cap ? reinterpret_cast<OrgS<MAXC> *>(W.cp_state + (size_t)q * COUPLED_STATE_BYTES) : nullptr; };
    for (int q = gwarp; q < nbig; q += gwarps) {
        Ctx c(s, w, K, W.big_list[q]);
        c.load(); c.prestep_pairs(); c.schedule(); c.template prestep_contacts<true>();
        if (dt_coef == 0.f) c.zero_pin_acc();                  // no warm start will consume the cached pin impulses
        c.store_contacts(W.ic_bounce, W.ic_jbias);
        if (OrgS<MAXC> *g = parked(q)) c.save_static(g);
        if (lane_id() == 0) atomicAdd(&W.flags[FL_NIC], s.nic);
        __syncwarp();
    }
    sync();
    // one pass of a coupled organism's own arbiters and constraints between two rounds of cross contacts
    auto org_pass = [&](int q, bool warm, bool last) {
        Ctx c(s, w, K, W.big_list[q]);
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
        for (int q = gwarp; q < nbig; q += gwarps) org_pass(q, true, false);
        sync();
    }
    const int iterations = w.p.iterations;
    for (int it = 0; it < iterations; it++) {
        for (int L = 1; L <= maxL; L++) { cross_level_pass<false>(w, W, L, nx, 0.f, gtid, gthreads); sync(); }
        for (int q = gwarp; q < nbig; q += gwarps) org_pass(q, false, it == iterations - 1);
        sync();
    }
}

struct GridSync { cg::grid_group g; __device__ void operator()() { g.sync(); } };
struct ClusterSync { cg::cluster_group c; __device__ void operator()() { __threadfence(); c.sync(); } };

// cooperative launch over a grid of CTAs: crowds (a component of dozens to hundreds of organisms), and the form the tests force
__global__ void __launch_bounds__(SIMPLE_WARPS * 32) k_coupled_simple(WORLD_ARG, Workspace W, StepK K) {
    if (W.flags[FL_NX] == 0 || !W.flags[FL_COMP_OVERFLOW] || W.flags[FL_BIG_NORG] <= CLUSTER_ORGS) return;   // uniform over the grid: no barrier has been entered
    GridSync sync{cg::this_grid()};
    coupled_body(w, W, K, sync);
}
// one thread-block cluster: the oversize components of an ordinary tick
__global__ void __cluster_dims__(CLUSTER_CTAS, 1, 1) __launch_bounds__(SIMPLE_WARPS * 32) k_coupled_cluster(WORLD_ARG, Workspace W, StepK K) {
    if (W.flags[FL_NX] == 0 || !W.flags[FL_COMP_OVERFLOW] || W.flags[FL_BIG_NORG] > CLUSTER_ORGS) return;
    ClusterSync sync{cg::this_cluster()};
    coupled_body(w, W, K, sync);
}
