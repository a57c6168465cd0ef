// components.cuh -- organisms and dead bodies that share cross contacts, solved one connected component per CTA.
//
// The cross-contact graph of a crowded world falls apart into thousands of small components (tick 3000 of the 100k-organism
// bench world: 27k components, none with more than 16 organisms; half of the coupled organisms touch dead bodies only). A
// component is independent of every other one, so its whole solve -- pre-step, exact warm start, ten sweeps of [cross contacts
// in canonical order, then each organism's own intra contacts and pin joints] -- runs inside one CTA with everything resident
// in shared memory: one warp per organism (the organism solver of orgsolve.cuh on a slab sized by the organism), the dead
// bodies and the cross contacts of the component in CTA-level tables. No state is parked in global memory and there is no grid
// barrier; the only global traffic is the load and the final store.
//
// Planning (a handful of small kernels, no host involvement):
//   k_comp_init     touched nodes (organism o -> node o, dead body d -> node n_org + d) become their own parents, counters zeroed
//   k_comp_union    lock-free union-find over the cross contacts (hook the larger root under the smaller: root == min node)
//   k_comp_count    root of every contact / coupled organism / coupled dead body; per-root counts and shared-memory need
//   k_comp_assign   every root picks the smallest CTA tier that holds it and reserves its run in the contact pool
//   k_comp_scatter  contacts into their component's run (the CTA sorts its run: canonical order == ascending contact index)
// A component that exceeds the largest tier is marked (cn_tier = -1) and solved by the cooperative kernel of coupled.cuh, which is
// launched every tick and returns at once when no component is marked (decided on the device: no host round trip).
#pragma once
#include "orgsolve.cuh"

// ---- tiers: warps (organisms) per CTA, shared memory, table capacities -------------------------------------------------------
#define COMP_TIERS 7
struct CompTier { int warps; int smem; int xmax; int dmax; };
// A warp is an organism: the three one-warp tiers take one organism of up to 24 / 40 / 64 living cells with its debris (half of the
// coupled organisms touch dead bodies only), so that no CTA carries a warp without work.
__host__ __device__ inline CompTier comp_tier(int t) {
    switch (t) {
        case 0: return CompTier{1, 12 * 1024, 24, 24};          // a small organism with its debris, or debris alone: 18 CTAs per SM
        case 1: return CompTier{1, 25 * 1024, 40, 40};          // one organism of up to 40 cells: 8 per SM
        case 2: return CompTier{1, 51 * 1024, 64, 48};          // one organism of any size: 4 per SM
        case 3: return CompTier{2, 55 * 1024, 96, 64};          // two organisms: 4 per SM
        case 4: return CompTier{4, 110 * 1024, 192, 128};       // up to four organisms: 2 per SM
        case 5: return CompTier{16, 220 * 1024, 384, 192};      // up to sixteen: 1 per SM
        default: return CompTier{16, 227 * 1024 - 512, 48, 32};     // a handful of LARGE organisms touching each other (5-8 x 27-44 KB of slabs, few contacts):
                                                                // small tables leave 221 KB for the slabs -- what used to fall back to the cooperative kernel on most ticks
    }
}
// slab class of an organism by its cell rows
__host__ __device__ inline int comp_slab_class(int nc) { return nc <= 24 ? 0 : (nc <= 40 ? 1 : (nc <= 48 ? 2 : 3)); }
__host__ __device__ inline int comp_slab_bytes(int k) {
    return k == 0 ? (int)((sizeof(OrgS<24>) + 15) & ~15u) : k == 1 ? (int)((sizeof(OrgS<40>) + 15) & ~15u) : k == 2 ? (int)((sizeof(OrgS<48>) + 15) & ~15u) : (int)((sizeof(OrgS<64>) + 15) & ~15u);
}

// one cross contact of the component, resident for the whole solve; bodies are addressed by their byte offset in shared memory
struct alignas(16) XCon {
    uint32_t vmA, bbA, vmB, bbB;                 // float4 (vx, vy, w, 1/m) and float4 (bias vx, bias vy, r/I, r) of the two bodies
    float nx, ny, bias, bounce;
    float nMass, tMass, mu, pad0;                // effective masses: constant for the whole solve, computed once when the record is built
    float jn, jt, jb, pad1;
};
__host__ __device__ inline int comp_table_bytes(const CompTier &T) {
    // contacts, dead bodies (two float4 each) + wall entry, sorted ids, contact order / steps, per-organism row -> rank maps, header
    return (15 + T.xmax * (int)sizeof(XCon) + (T.dmax + 1) * 32 + (2 * T.xmax + 16) * 4 /* endpoint ids */ + T.xmax * 4 /* global index */ +
           T.xmax * 2 /* order */ + (T.xmax + 2) * 2 /* step starts */ + T.xmax * 2 /* levels */ + T.warps * 64 /* rank_of */ + T.dmax * 4 /* dead ids */ + 256) & ~15;
}

// ---- planning ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int comp_find(int *parent, int v) {
    volatile int *P = parent;
    for (;;) {
        int p = P[v];
        if (p == v) return v;
        int gp = P[p];
        if (gp == p) return p;
        P[v] = gp;                                       // path halving (benign race: any ancestor is a valid parent)
        v = gp;
    }
}
__global__ void __launch_bounds__(256) k_comp_init(WORLD_ARG, Workspace W) {
    const int NC = w.p.n_org * MAXC, NB = NC + w.p.n_dead, nx = W.flags[FL_NX];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nx; c += gridDim.x * blockDim.x) {
        const uint64_t key = W.nx_key[c];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int v = comp_node(e ? (uint32_t)key : (uint32_t)(key >> 32), NC, NB, w.p.n_org);
            if (v < 0) continue;
            W.cn_parent[v] = v; W.cn_nx[v] = 0; W.cn_norg[v] = 0; W.cn_ndead[v] = 0; W.cn_need[v] = 0; W.cn_fill[v] = 0; W.cn_seen[v] = 0;
        }
    }
}
__global__ void __launch_bounds__(256) k_comp_union(WORLD_ARG, Workspace W) {
    const int NC = w.p.n_org * MAXC, NB = NC + w.p.n_dead, nx = W.flags[FL_NX];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nx; c += gridDim.x * blockDim.x) {
        const uint64_t key = W.nx_key[c];
        int a = comp_node((uint32_t)(key >> 32), NC, NB, w.p.n_org), b = comp_node((uint32_t)key, NC, NB, w.p.n_org);
        if (a < 0 || b < 0 || a == b) continue;
        for (;;) {
            int ra = comp_find(W.cn_parent, a), rb = comp_find(W.cn_parent, b);
            if (ra == rb) break;
            if (ra < rb) { int t = ra; ra = rb; rb = t; }
            if (atomicCAS(&W.cn_parent[ra], ra, rb) == ra) break;             // the larger root goes under the smaller one
        }
    }
}
// per contact: its component; per coupled body (once: the first to flip cn_seen): the component's organism / dead counts and need
__global__ void __launch_bounds__(256) k_comp_count(WORLD_ARG, Workspace W) {
    const int NC = w.p.n_org * MAXC, NB = NC + w.p.n_dead, nx = W.flags[FL_NX], n_org = w.p.n_org;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nx; c += gridDim.x * blockDim.x) {
        const uint64_t key = W.nx_key[c];
        const int a = comp_node((uint32_t)(key >> 32), NC, NB, n_org);                 // body a is never a wall
        const int root = comp_find(W.cn_parent, a);
        W.comp_of[c] = root;
        atomicAdd(&W.cn_nx[root], 1);
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int v = comp_node(e ? (uint32_t)key : (uint32_t)(key >> 32), NC, NB, n_org);
            if (v < 0 || atomicExch(&W.cn_seen[v], 1)) continue;
            if (v < n_org) { atomicAdd(&W.cn_norg[root], 1); atomicAdd(&W.cn_need[root], comp_slab_bytes(comp_slab_class(W.org_nlive[v]))); }
            else atomicAdd(&W.cn_ndead[root], 1);
        }
    }
}
__global__ void __launch_bounds__(256) k_comp_assign(WORLD_ARG, Workspace W, int force_fallback) {
    const int NC = w.p.n_org * MAXC, NB = NC + w.p.n_dead, nx = W.flags[FL_NX], n_org = w.p.n_org;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nx; c += gridDim.x * blockDim.x) {
        const int root = W.comp_of[c];
        if (atomicExch(&W.cn_seen[root], 2) == 2) continue;                            // one thread per component
        // (cn_seen of a root is 1 after k_comp_count: every root is an endpoint of one of its contacts)
        const int cnt = W.cn_nx[root], no = W.cn_norg[root], nd = W.cn_ndead[root], need = W.cn_need[root];
        int tier = -1;
        for (int t = 0; t < COMP_TIERS; t++) {
            const CompTier T = comp_tier(t);
            if (no <= T.warps && cnt <= T.xmax && nd <= T.dmax && need + comp_table_bytes(T) <= T.smem) { tier = t; break; }
        }
        if (force_fallback) tier = -1;
        W.cn_tier[root] = tier;
#ifdef COMP_DIAG
        if (tier < 0) printf("OVERSIZE tick %d organisms %d contacts %d dead %d slab bytes %d\n", (int)w.g_tick[0], no, cnt, nd, need);
#endif
        if (tier < 0) { W.flags[FL_COMP_OVERFLOW] = 1; atomicAdd(&W.flags[FL_BIG_NORG], no); continue; }   // this component goes to the fallback kernel (coupled.cuh)
        W.cn_xoff[root] = atomicAdd(&W.flags[FL_COMP_POOL], cnt);
        const int q = atomicAdd(&W.flags[FL_COMP_N0 + tier], 1);
        W.comp_list[(size_t)tier * 2 * w.p.x_cap + q] = root;
    }
    (void)NC; (void)NB; (void)n_org;
}
__global__ void __launch_bounds__(256) k_comp_scatter(WORLD_ARG, Workspace W) {
    const int nx = W.flags[FL_NX];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nx; c += gridDim.x * blockDim.x) {
        const int root = W.comp_of[c];
        if (W.cn_tier[root] < 0) continue;
        W.comp_pool[W.cn_xoff[root] + atomicAdd(&W.cn_fill[root], 1)] = c;
    }
}

// ---- the solve -----------------------------------------------------------------------------------------------------------------
struct CompS {                                   // CTA header, at the start of the table region
    int nx, norg, ndead, nsteps, maxlevel, overflow;
    int org[16]; int slab_off[16]; int slab_cls[16];
};
__device__ __forceinline__ void comp_sync() { __syncthreads(); }

// cpArbiterApplyImpulse on two bodies resident in shared memory (the arithmetic of contact_apply in steptypes.cuh; the effective
// masses come from the record; accumulate-and-clamp as contact_accumulate, like the intra contacts of OrgCtx::contacts_pass)
__device__ __forceinline__ void xcon_apply(unsigned char *sm, XCon &c) {
    float4 va = *reinterpret_cast<float4 *>(sm + c.vmA), ba = *reinterpret_cast<float4 *>(sm + c.bbA);
    float4 vb = *reinterpret_cast<float4 *>(sm + c.vmB), bb = *reinterpret_cast<float4 *>(sm + c.bbB);
    const float4 g = *reinterpret_cast<const float4 *>(&c.nx), k = *reinterpret_cast<const float4 *>(&c.nMass), ac = *reinterpret_cast<const float4 *>(&c.jn);
    const float nx = g.x, ny = g.y, nMass = k.x, tMass = k.y, mu = k.z;
    const float dvx = vb.x - va.x, dvy = vb.y - va.y;
    const float vbn = (bb.x - ba.x) * nx + (bb.y - ba.y) * ny;
    const float vrn = dvx * nx + dvy * ny;
    const float vrt = (dvy * nx - dvx * ny) - bb.w * vb.z - ba.w * va.z;
    const float jbnOld = ac.z, jBias = fmaxf(jbnOld + (g.z - vbn) * nMass, 0.0f);
    float dn, dtg, jnS, jtS;
    contact_accumulate(ac.x, ac.y, -(g.w + vrn) * nMass, -vrt * tMass, mu, jnS, jtS, dn, dtg);
    *reinterpret_cast<float4 *>(&c.jn) = make_float4(jnS, jtS, jBias, 0.f);
    const float bj = jBias - jbnOld;
    ba.x -= nx * bj * va.w; ba.y -= ny * bj * va.w;
    bb.x += nx * bj * vb.w; bb.y += ny * bj * vb.w;
    const float jx = nx * dn - ny * dtg, jy = nx * dtg + ny * dn;
    va.x -= jx * va.w; va.y -= jy * va.w; va.z -= ba.z * dtg;
    vb.x += jx * vb.w; vb.y += jy * vb.w; vb.z -= bb.z * dtg;
    *reinterpret_cast<float4 *>(sm + c.vmA) = va; *reinterpret_cast<float2 *>(sm + c.bbA) = make_float2(ba.x, ba.y);
    *reinterpret_cast<float4 *>(sm + c.vmB) = vb; *reinterpret_cast<float2 *>(sm + c.bbB) = make_float2(bb.x, bb.y);
}

struct CompView {                                // where the CTA's tables live (byte offsets computed once from the tier)
    unsigned char *sm; CompS *hd; XCon *xc; float4 *dead; uint32_t *ep; int *gidx; uint16_t *order, *sstart, *level; uint8_t *rank_of; int *dead_id;
    uint32_t wall_vm, wall_bb;
};
__device__ __forceinline__ CompView comp_view(unsigned char *sm, int slab_bytes_total, const CompTier &T) {
    CompView V; V.sm = sm;
    unsigned char *p = sm + slab_bytes_total;
    V.hd = reinterpret_cast<CompS *>(p); p += 256;
    V.xc = reinterpret_cast<XCon *>(p); p += T.xmax * sizeof(XCon);
    V.dead = reinterpret_cast<float4 *>(p); p += (T.dmax + 1) * 32;
    V.wall_vm = (uint32_t)((unsigned char *)(V.dead + 2 * T.dmax) - sm); V.wall_bb = V.wall_vm + 16;
    V.ep = reinterpret_cast<uint32_t *>(p); p += (2 * T.xmax + 16) * 4;
    V.gidx = reinterpret_cast<int *>(p); p += T.xmax * 4;
    V.dead_id = reinterpret_cast<int *>(p); p += T.dmax * 4;
    V.order = reinterpret_cast<uint16_t *>(p); p += T.xmax * 2;
    V.sstart = reinterpret_cast<uint16_t *>(p); p += (T.xmax + 2) * 2;
    V.level = reinterpret_cast<uint16_t *>(p); p += T.xmax * 2;
    V.rank_of = p;
    return V;
}

// one sweep over the component's cross contacts in canonical order: a level step at a time, all threads of the CTA
__device__ __forceinline__ void comp_cross_pass(const CompView &V) {
    const int ns = V.hd->nsteps;
    for (int st = 0; st < ns; st++) {
        const int pos = V.sstart[st] + threadIdx.x;
        if (pos < V.sstart[st + 1]) xcon_apply(V.sm, V.xc[V.order[pos]]);
        comp_sync();
    }
}

// the organism of a warp, phase by phase (the CTA barriers sit between the phases, in k_component): the context is rebuilt from the
// slab, which stays resident for the whole solve
template <int CM> __device__ __forceinline__ OrgS<CM> &comp_slab(const CompView &V, int wslot) { return *reinterpret_cast<OrgS<CM> *>(V.sm + V.hd->slab_off[wslot]); }
template <int CM>
__device__ void comp_org_load(const SapioWorld &w, const Workspace &W, const StepK &K, const CompView &V, int o, int wslot) {
    OrgS<CM> &s = comp_slab<CM>(V, wslot);
    OrgCtx<CM, 32> c(s, w, K, o);
    c.load(); c.prestep_pairs(); c.schedule(); c.template prestep_contacts<true>();
    uint8_t *ro = V.rank_of + wslot * 64;                              // row -> rank of the living cells, for the contacts' body references
    const int lane = lane_id();
    ro[lane] = 255; ro[lane + 32] = 255;
    __syncwarp();
    if (lane < c.n) ro[s.row_of[lane]] = (uint8_t)lane;
    if (lane + 32 < c.n) ro[s.row_of[lane + 32]] = (uint8_t)(lane + 32);
    if (lane == 0) atomicAdd(&W.flags[FL_NIC], s.nic);
}
template <int CM>
__device__ void comp_org_warm(const SapioWorld &w, const Workspace &W, const StepK &K, const CompView &V, int o, int wslot, float dt_coef) {
    OrgCtx<CM, 32> c(comp_slab<CM>(V, wslot), w, K, o);
    c.resume();
    if (dt_coef != 0.f) { const CrossView X = cross_view(W); c.warm_start(dt_coef, &X); }
    else c.zero_pin_acc();
}
template <int CM>
__device__ void comp_org_sweep(const SapioWorld &w, const StepK &K, const CompView &V, int o, int wslot) {
    OrgS<CM> &s = comp_slab<CM>(V, wslot);
    OrgCtx<CM, 32> c(s, w, K, o);
    c.resume();
    c.template contacts_pass<false>(0.f, s.nsteps);
    c.template joints_pass<false>(0.f, 2 * c.n - 3);
}
template <int CM>
__device__ void comp_org_store(const SapioWorld &w, const StepK &K, const CompView &V, int o, int wslot) {
    OrgCtx<CM, 32> c(comp_slab<CM>(V, wslot), w, K, o);
    c.resume();
    c.store_vel(); c.store_contacts(nullptr, nullptr);
}
#define COMP_DISPATCH(cls, fn, ...) do { if ((cls) == 0) fn<24>(__VA_ARGS__); else if ((cls) == 1) fn<40>(__VA_ARGS__); else if ((cls) == 2) fn<48>(__VA_ARGS__); else fn<64>(__VA_ARGS__); } while (0)

// TW: warps of the tier's CTA (one instantiation per CTA size, so that the small tiers are not compiled under the 16-warp register cap)
template <int TW>
__global__ void __launch_bounds__(TW * 32, TW == 1 ? 16 : (TW == 2 ? 5 : (TW == 4 ? 2 : 1))) k_component(WORLD_ARG, Workspace W, StepK K, int tier) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const CompTier T = comp_tier(tier);
    const int NC = w.p.n_org * MAXC, NB = NC + w.p.n_dead;
    const int n_comp = W.flags[FL_COMP_N0 + tier];
    const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31;
    const int slab_total = T.smem - comp_table_bytes(T);
    const CompView V = comp_view(smem_raw, slab_total, T);
    const float dt_coef = w.g_stepped[0] ? 1.0f : 0.0f;
    const int iterations = w.p.iterations;
    __shared__ int sh_q;
    for (;;) {
        comp_sync();
        if (tid == 0) sh_q = atomicAdd(&W.flags[FL_COMP_WORK0 + tier], 1);
        comp_sync();
        const int q = sh_q;
        if (q >= n_comp) break;
        const int root = W.comp_list[(size_t)tier * 2 * w.p.x_cap + q];
        const int nx = min(W.cn_nx[root], T.xmax), xoff = W.cn_xoff[root];
        // ---- gather: the component's contacts in canonical order (ascending global index) ------------------------------------------
        for (int i = tid; i < nx; i += nthr) {
            const int me = W.comp_pool[xoff + i];
            int r = 0;
            for (int j = 0; j < nx; j++) r += W.comp_pool[xoff + j] < me;
            V.gidx[r] = me;
        }
        comp_sync();
        const int ne = 2 * nx;
        for (int i = tid; i < ne; i += nthr) {                                    // endpoints; walls are nobody
            const uint64_t key = W.nx_key[V.gidx[i >> 1]];
            const uint32_t b = (i & 1) ? (uint32_t)key : (uint32_t)(key >> 32);
            V.ep[i] = (int)b < NB ? b : 0xFFFFFFFFu;
        }
        comp_sync();
        if (warp == 0) {                                                          // distinct organisms (<= 16) and dead bodies of the component
            int norg = 0, ndead = 0;
            for (int i = 0; i < ne; i++) {
                const uint32_t v = V.ep[i];
                if (v == 0xFFFFFFFFu) continue;
                if ((int)v < NC) {
                    const int o = (int)(v >> 6);
                    const bool hit = lane < min(norg, 16) && V.hd->org[lane] == o;
                    if (!__any_sync(FULL, hit)) { if (lane == 0 && norg < 16) V.hd->org[norg] = o; norg++; }
                } else {
                    const int d = (int)v - NC;
                    bool hit = false;
                    for (int j = lane; j < min(ndead, T.dmax); j += 32) hit |= V.dead_id[j] == d;
                    if (!__any_sync(FULL, hit)) { if (lane == 0 && ndead < T.dmax) V.dead_id[ndead] = d; ndead++; }
                }
                __syncwarp();
            }
            if (lane == 0) {
                bool over = norg > T.warps || norg > 16 || ndead > T.dmax || W.cn_nx[root] > T.xmax;
                const int no = min(norg, 16);
                for (int i = 1; i < no; i++) { int v = V.hd->org[i], j = i - 1; while (j >= 0 && V.hd->org[j] > v) { V.hd->org[j + 1] = V.hd->org[j]; j--; } V.hd->org[j + 1] = v; }
                int off = 0;
                for (int i = 0; i < no; i++) {
                    const int cls = comp_slab_class(W.org_nlive[V.hd->org[i]]);
                    V.hd->slab_cls[i] = cls; V.hd->slab_off[i] = off; off += comp_slab_bytes(cls);
                }
                if (off > slab_total) over = true;
                V.hd->nx = nx; V.hd->norg = no; V.hd->ndead = min(ndead, T.dmax); V.hd->overflow = over;
                if (over) { W.flags[FL_OVERFLOW] = 1; atomicAdd((unsigned long long *)&w.g_stats[SAPIO_STAT_OVERFLOW], 1ull); }   // the planner sized this component wrongly: never silent
            }
        }
        comp_sync();
        if (V.hd->overflow) continue;                                             // (uniform) leave it untouched rather than corrupt memory
        const int norg = V.hd->norg, ndead = V.hd->ndead;
        const bool has_org = warp < norg;
        const int my_org = has_org ? V.hd->org[warp] : -1, my_cls = has_org ? V.hd->slab_cls[warp] : 0;
        // ---- phase A: organisms load and pre-step (one warp each); dead bodies into the table (pre-warm-start state) ----------------
        if (has_org) COMP_DISPATCH(my_cls, comp_org_load, w, W, K, V, my_org, warp);
        for (int i = tid; i <= ndead; i += nthr) {
            if (i == ndead) { V.dead[2 * T.dmax] = make_float4(0.f, 0.f, 0.f, 0.f); V.dead[2 * T.dmax + 1] = make_float4(0.f, 0.f, 0.f, w.p.frame_width); continue; }   // wall
            const int d = V.dead_id[i];
            const float2 v = reinterpret_cast<const float2 *>(w.d_vel)[d], vb = reinterpret_cast<const float2 *>(w.d_vbias)[d];
            const float r = w.d_size[d] * 0.5f, minv = 1.0f / w.d_mass[d], iinv = 1.0f / (0.5f * w.d_mass0[d] * r * r);
            V.dead[2 * i] = make_float4(v.x, v.y, w.d_angvel[d], minv);
            V.dead[2 * i + 1] = make_float4(vb.x, vb.y, iinv * r, r);
        }
        comp_sync();
        // ---- contact records: where the two bodies live, constants and cached impulses (k_cross_prestep) ----------------------------
        for (int i = tid; i < nx; i += nthr) {
            const int g = V.gidx[i];
            const uint64_t key = W.nx_key[g];
            XCon x;
            float ua = 0.f, ub = 0.f;
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const uint32_t b = e ? (uint32_t)key : (uint32_t)(key >> 32);
                uint32_t vm, bb; float u;
                if ((int)b < NC) {
                    const int o = (int)(b >> 6);
                    int k = 0; while (k < norg - 1 && V.hd->org[k] != o) k++;
                    const int rank = V.rank_of[k * 64 + (b & 63)], cls = V.hd->slab_cls[k];
                    const size_t ovm = cls == 0 ? offsetof(OrgS<24>, vm) : cls == 1 ? offsetof(OrgS<40>, vm) : cls == 2 ? offsetof(OrgS<48>, vm) : offsetof(OrgS<64>, vm);
                    const size_t obb = cls == 0 ? offsetof(OrgS<24>, bb) : cls == 1 ? offsetof(OrgS<40>, bb) : cls == 2 ? offsetof(OrgS<48>, bb) : offsetof(OrgS<64>, bb);
                    vm = (uint32_t)(V.hd->slab_off[k] + ovm + 16 * rank); bb = (uint32_t)(V.hd->slab_off[k] + obb + 16 * rank);
                    u = w.c_fric[b];
                } else if ((int)b < NB) {
                    const int d = (int)b - NC;
                    int k = 0; while (k < ndead - 1 && V.dead_id[k] != d) k++;
                    vm = (uint32_t)((unsigned char *)(V.dead + 2 * k) - V.sm); bb = vm + 16;
                    u = w.d_fric[d];
                } else { vm = V.wall_vm; bb = V.wall_bb; u = w.p.wall_fric; }
                if (e) { x.vmB = vm; x.bbB = bb; ub = u; } else { x.vmA = vm; x.bbA = bb; ua = u; }
            }
            const float2 n = W.nx_n[g];
            x.nx = n.x; x.ny = n.y; x.bias = W.nx_bias[g]; x.bounce = W.nx_bounce[g]; x.mu = ua * ub;
            {   // k_n = 1/ma + 1/mb, k_t = k_n + ra^2/Ia + rb^2/Ib (circles: r x n = 0); a wall has 1/m = r/I = 0
                const float4 va = *reinterpret_cast<const float4 *>(V.sm + x.vmA), ba = *reinterpret_cast<const float4 *>(V.sm + x.bbA);
                const float4 vb = *reinterpret_cast<const float4 *>(V.sm + x.vmB), bb = *reinterpret_cast<const float4 *>(V.sm + x.bbB);
                x.nMass = 1.0f / (va.w + vb.w);
                x.tMass = 1.0f / (va.w + vb.w + ba.z * ba.w + bb.z * bb.w);
            }
            x.pad0 = 0.f; x.pad1 = 0.f;
            x.jn = W.nx_jn[g]; x.jt = W.nx_jt[g]; x.jb = 0.f;
            V.xc[i] = x;
        }
        comp_sync();
        // ---- dependency levels of the contacts (contact i runs after every earlier one that shares a body with it; walls are nobody):
        // the exact sequential order, evaluated a level at a time. Thread 0 schedules while the rest warm-start the dead bodies.
        if (tid == 0) {
            int mx = 0;
            for (int i = 0; i < nx; i++) {
                const uint32_t a = V.xc[i].vmA, b = V.xc[i].vmB;
                int L = 0;
                for (int j = i - 1; j >= 0; j--) {                                 // latest earlier contact on either body
                    const uint32_t ja = V.xc[j].vmA, jb = V.xc[j].vmB;
                    if (ja == a || jb == a || (b != V.wall_vm && (ja == b || jb == b))) L = max(L, (int)V.level[j]);
                    if (L == mx) break;                                            // cannot grow further
                }
                L++;
                V.level[i] = (uint16_t)L; mx = max(mx, L);
            }
            // stable counting sort by level, steps of at most nthr contacts
            int ns = 0, pos = 0;
            for (int L = 1; L <= mx; L++) {
                const int start = pos;
                for (int i = 0; i < nx; i++) if (V.level[i] == L) V.order[pos++] = (uint16_t)i;
                for (int k = start; k < pos; k += nthr) V.sstart[ns++] = (uint16_t)k;
            }
            V.sstart[ns] = (uint16_t)nx;
            V.hd->nsteps = ns; V.hd->maxlevel = mx;
            atomicMax(&W.flags[FL_MAXLEVEL], mx);
        }
        if (dt_coef != 0.f) {                                                     // exact warm start of the dead bodies (cross_warm_dead of coupled.cuh, on the table)
            const int nxg = W.flags[FL_NX];
            const double dt = (double)dt_coef;
            for (int i = tid; i < ndead; i += nthr) {
                const uint32_t body = (uint32_t)(NC + V.dead_id[i]);
                double ax = 0.0, ay = 0.0, at = 0.0;
                for (int qq = W.off_all[body], e = W.off_all[body + 1]; qq < e; qq++) {
                    const int c = W.adj_c[qq];
                    if (c < 0 || c >= nxg) continue;
                    const float2 nn = W.nx_n[c];
                    const double jn = (double)W.nx_jn[c] * dt, jt = (double)W.nx_jt[c] * dt, nxx = (double)nn.x, nyy = (double)nn.y;
                    const double jx = nxx * jn - nyy * jt, jy = nxx * jt + nyy * jn;
                    if ((uint32_t)(W.nx_key[c] >> 32) == body) { ax -= jx; ay -= jy; } else { ax += jx; ay += jy; }
                    at += jt;
                }
                float4 vm = V.dead[2 * i]; const float4 bb = V.dead[2 * i + 1];
                vm.x = (float)fma(ax, (double)vm.w, (double)vm.x); vm.y = (float)fma(ay, (double)vm.w, (double)vm.y);
                vm.z = (float)((double)vm.z - (double)bb.z * at);
                V.dead[2 * i] = vm;
            }
        }
        if (has_org) COMP_DISPATCH(my_cls, comp_org_warm, w, W, K, V, my_org, warp, dt_coef);
        comp_sync();
        // ---- ten sweeps: cross contacts in canonical order, then every organism's own arbiters and constraints ---------------------
        for (int it = 0; it < iterations; it++) {
            comp_cross_pass(V);
            if (has_org) COMP_DISPATCH(my_cls, comp_org_sweep, w, K, V, my_org, warp);
            comp_sync();
        }
        // ---- store ------------------------------------------------------------------------------------------------------------------
        if (has_org) COMP_DISPATCH(my_cls, comp_org_store, w, K, V, my_org, warp);
        for (int i = tid; i < ndead; i += nthr) {
            const int d = V.dead_id[i];
            const float4 vm = V.dead[2 * i], bb = V.dead[2 * i + 1];
            reinterpret_cast<float2 *>(w.d_vel)[d] = make_float2(vm.x, vm.y); w.d_angvel[d] = vm.z;
            reinterpret_cast<float2 *>(w.d_vbias)[d] = make_float2(bb.x, bb.y);
        }
        for (int i = tid; i < nx; i += nthr) { const int g = V.gidx[i]; W.nx_jn[g] = V.xc[i].jn; W.nx_jt[g] = V.xc[i].jt; }
    }
}
