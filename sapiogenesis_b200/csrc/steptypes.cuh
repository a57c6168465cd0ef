// steptypes.cuh -- what the translation units of Space.step share: the workspace layout, the step constants and the contact
// arithmetic (device inline). No kernels here.
#pragma once
#include <cub/cub.cuh>
#include "common.cuh"
// ------------------------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------------------------
struct Workspace {
    int *pop;                       // population reduction scratch
    int *flags;                     // FL_*
    uint32_t *gkey, *gval, *skey, *sval;   // grid: cell key of every body slot [NB], list of living bodies, (cell key, body) in cell order [living]
    uint32_t *stmp; float4 *srec;   // bodies in cell order before the ranking pass; (x, y, radius, id bits) in cell order (grid.cuh)
    int *cell_cnt, *scan_parts;     // [grid cells + 1] bodies per cell (zero between builds); tile sums of the scans
    void *cub_tmp; size_t cub_bytes;
    int *cell_start;                // [grid cells + 1] run of cell c in the cell-ordered lists: [cell_start[c], cell_start[c + 1])
    int *cnt_up, *cnt_all;          // per body counts [NB + 1]
    int *off_up, *off_all;          // exclusive scans [NB + 1]
    uint64_t *nx_key;               // new contact list [x_cap]
    float *nx_jn, *nx_jt, *nx_bounce, *nx_jbias, *nx_bias;
    float2 *nx_n;                   // contact normal (a -> b)
    int *nx_level, *nx_prev_a, *nx_prev_b;
    uint32_t *adj_other;            // adjacency: partner body id per entry [2*x_cap]
    int *adj_c;                     // contact index of every adjacency entry [2*x_cap]
    uint32_t *adj_tmp;              // partner runs in discovery order [2*x_cap]
    int2 *hit_list;                 // (body, start of its run) [2*x_cap]
    uint8_t *org_coupled;           // [n_org]
    uint8_t *org_nlive;             // [n_org] living cells of every organism (k_classify_count): what sizes its solver slab
    int *coupled_list;              // [n_org]
    int *sorted_orgs;               // [n_org] uncoupled organisms sorted by cell rows
    int *size_hist;                 // [MAXC + 1] inside the flags block: histogram, then scatter cursors
    float *ic_bounce, *ic_jbias;    // [n_org * ICAP] (coupled path only)
    float4 *pair_cst;               // [4 classes][ORG_GC_WARPS][NPAIR] pin constants of the organism a warp is solving (GC classes)
    // connected components of the cross-contact graph (components.cuh): per node (organism o -> o, dead body d -> n_org + d)
    int *cn_parent, *cn_nx, *cn_norg, *cn_ndead, *cn_need, *cn_xoff, *cn_fill, *cn_seen, *cn_tier;
    int *comp_of, *comp_pool;      // [x_cap] component (root node) of a contact; contacts grouped by component
    int *comp_list;                 // [COMP_TIERS][2 * x_cap] roots by CTA tier
    int *big_list;                  // [n_org] organisms of the oversize components, compacted by the fallback kernel itself
    char *cp_state; int cp_cap;     // parked solver state of the first cp_cap coupled organisms (sizeof(OrgS<MAXC>) each)
};
enum { FL_NX = 0, FL_NADJ = 1, FL_OVERFLOW = 2, FL_MAXLEVEL = 3, FL_NCOUPLED = 5, FL_NIC = 6, FL_CLS0 = 8 /* ..16: class starts + end */,
       FL_WORK0 = 20 /* ..27 */, FL_TMPTOP = 28, FL_NHIT = 29, FL_RARE = 30, FL_NLIVE = 31 /* living bodies in the grid */, FL_HIST = 32 /* ..96 */, FL_CHIST = 100 /* ..164: coupled organisms by rows */,
       FL_CCLS0 = 168 /* ..176: class starts + end in coupled_list */, FL_CHG0 = 180 /* ..182 */, FL_CWORK0 = 200 /* ..203: two alternating pairs of work counters of k_coupled */,
       FL_COMP_OVERFLOW = 210 /* a component exceeds the largest CTA tier: the tick falls back to the cooperative kernel */, FL_COMP_POOL = 211, FL_BIGLEVEL = 209,
       FL_COMP_N0 = 212 /* ..218: components per tier */, FL_COMP_WORK0 = 220 /* ..226: work counters */,
       FL_BIG_NORG = 230 /* organisms in components beyond the largest CTA tier */, FL_BIG_N = 231 /* length of big_list */, FL_BYTES = 1024 };

#define COUPLED_STATE_BYTES 49152      /* >= sizeof(OrgS<MAXC>), checked where the struct is known */
// sort only the bits a grid key can have; the key of an empty slot (all ones) still sorts behind every cell: its low bits are all ones
static inline int grid_key_bits(int ncell) { int b = 1; while (b < 32 && ((1ll << b) - 1) < (long long)ncell) b++; return b; }

static inline size_t carve_workspace(const SapioParams &p, char *base, Workspace *W) {
    size_t off = 0;
    auto take = [&](size_t bytes) -> char * { char *r = base ? base + off : nullptr; off += (bytes + 255) & ~size_t(255); return r; };
    const size_t NB = (size_t)p.n_org * MAXC + p.n_dead, XC = (size_t)p.x_cap, NG = (size_t)p.grid_nx * p.grid_ny;
    W->pop = (int *)take(256);
    W->flags = (int *)take(FL_BYTES);
    W->size_hist = W->flags ? W->flags + FL_HIST : nullptr;
    W->gkey = (uint32_t *)take(NB * 4); W->gval = (uint32_t *)take(NB * 4);
    W->skey = (uint32_t *)take(NB * 4); W->sval = (uint32_t *)take(NB * 4);
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int *)nullptr, (int *)nullptr, (int)NB + 1);
    W->cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
    W->cub_tmp = take(W->cub_bytes);
    W->cell_start = (int *)take((NG + 1) * 4); W->cell_cnt = (int *)take((NG + 1) * 4);
    W->stmp = (uint32_t *)take(NB * 4); W->srec = (float4 *)take(NB * 16);
    { const size_t big = (NG > NB ? NG : NB) + 1; W->scan_parts = (int *)take(((big + 2047) / 2048 + 1) * 4); }
    W->cnt_up = (int *)take((NB + 1) * 4); W->cnt_all = (int *)take((NB + 1) * 4);
    W->off_up = (int *)take((NB + 1) * 4); W->off_all = (int *)take((NB + 1) * 4);
    W->nx_key = (uint64_t *)take(XC * 8);
    W->nx_jn = (float *)take(XC * 4); W->nx_jt = (float *)take(XC * 4);
    W->nx_bounce = (float *)take(XC * 4); W->nx_jbias = (float *)take(XC * 4);
    W->nx_bias = (float *)take(XC * 4); W->nx_n = (float2 *)take(XC * 8);
    W->nx_level = (int *)take(XC * 4); W->nx_prev_a = (int *)take(XC * 4); W->nx_prev_b = (int *)take(XC * 4);
    W->adj_other = (uint32_t *)take(2 * XC * 4);
    W->adj_c = (int *)take(2 * XC * 4);
    W->adj_tmp = (uint32_t *)take(2 * XC * 4);
    W->hit_list = (int2 *)take(2 * XC * 8);
    W->org_coupled = (uint8_t *)take((size_t)p.n_org);
    W->org_nlive = (uint8_t *)take((size_t)p.n_org);
    W->coupled_list = (int *)take((size_t)p.n_org * 4);
    W->sorted_orgs = (int *)take((size_t)p.n_org * 4);
    W->ic_bounce = (float *)take((size_t)p.n_org * SAPIO_ICAP * 4);
    W->ic_jbias = (float *)take((size_t)p.n_org * SAPIO_ICAP * 4);
    W->pair_cst = (float4 *)take((size_t)4 * (148 * 32) * SAPIO_NPAIR * sizeof(float4));
    {
        const size_t NN = (size_t)p.n_org + p.n_dead;
        W->cn_parent = (int *)take(NN * 4); W->cn_nx = (int *)take(NN * 4); W->cn_norg = (int *)take(NN * 4); W->cn_ndead = (int *)take(NN * 4);
        W->cn_need = (int *)take(NN * 4); W->cn_xoff = (int *)take(NN * 4); W->cn_fill = (int *)take(NN * 4); W->cn_seen = (int *)take(NN * 4); W->cn_tier = (int *)take(NN * 4);
        W->comp_of = (int *)take(XC * 4); W->comp_pool = (int *)take(XC * 4); W->comp_list = (int *)take((size_t)7 * 2 * XC * 4);
    }
    W->big_list = (int *)take((size_t)p.n_org * 4);
    W->cp_cap = p.n_org < 8192 ? p.n_org : 8192;      // parked state of the fallback kernel (beyond: recomputed per pass)
    W->cp_state = take((size_t)W->cp_cap * COUPLED_STATE_BYTES);
    return off;
}

// ------------------------------------------------------------------------------------------------------------------
// contact arithmetic (cpArbiter.c) specialised for circles: every contact offset is r1 = ra*n, r2 = -rb*n (a wall
// counts as rb = frame radius on a static body), so r x n = 0 exactly. Writing that out removes the catastrophic
// cancellation the general r x j form has in fp32:
//   k(n) = ma + mb            k(t) = ma + mb + ia*ra^2 + ib*rb^2
//   v_rel.n = (vb - va).n     v_rel.t = (vb - va).t - rb*wb - ra*wa
//   torque of impulse (jn, jt): a: -ia*ra*jt   b: -ib*rb*jt      bias impulses (along n) never touch w_bias
// ------------------------------------------------------------------------------------------------------------------
struct StepK { double bc_contact, bc_joint, dt; int dbg; };      // dbg: A/B switches of diagnostic builds (SAPIO_DEBUG_FLAGS), 0 otherwise

struct CBody { float vx, vy, w, vbx, vby, minv, iinv, r; };

__device__ __forceinline__ float contact_bias(double d, float ra, float rb, const SapioParams &p, const StepK &K) {
    double dist = d - (double)ra - (double)rb;                            // ((r2 - r1) + (pb - pa)) . n
    return (float)(-K.bc_contact * fmin(0.0, dist + (double)p.collision_slop) / K.dt);
}
__device__ __forceinline__ float contact_vrn(const CBody &A, const CBody &B, float nx, float ny) { return (B.vx - A.vx) * nx + (B.vy - A.vy) * ny; }

// The accumulate-and-clamp of cpArbiterApplyImpulse,
//     jnAcc' = max(jnAcc + jn, 0)          jtAcc' = clamp(jtAcc + jt, -mu jnAcc', mu jnAcc')         apply (jnAcc' - jnAcc, jtAcc' - jtAcc),
// to the accuracy of a double-precision evaluation but in fp32 arithmetic: a resting contact squeezed between pinned cells carries 1e5 of
// impulse, and ulp(1e5) = 0.008 would quantise what a sweep adds if the DIFFERENCE of the rounded accumulators were applied. Unclamped, the
// difference is the increment itself, exactly; the sign of a rounded sum is the sign of the exact one; and the friction bound mu * jnAcc' is
// carried as an unevaluated sum of two floats (the product's and the sum's rounding errors from an FMA and a TwoSum), so the clamped branch
// loses nothing either. Replaces ~30 double-pipe instructions on the chain of every level step of an aged organism (most of its contacts
// carry more than the 256 below which round 1 stayed in fp32).
__device__ __forceinline__ void contact_accumulate(float jnO, float jtO, float jnI, float jtI, float mu, float &jnS, float &jtS, float &dn, float &dtg) {
    const float sN = jnO + jnI;
    const bool pos = sN > 0.0f;
    jnS = pos ? sN : 0.0f; dn = pos ? jnI : -jnO;
    const float bN = sN - jnO, eN = pos ? (jnO - (sN - bN)) + (jnI - bN) : 0.0f;          // jnAcc' == jnS + eN exactly
    const float mh = mu * jnS, ml = fmaf(mu, jnS, -mh) + mu * eN;                          // mu * jnAcc' == mh + ml to ~1e-14
    const float sT = jtO + jtI, bT = sT - jtO, eT = (jtO - (sT - bT)) + (jtI - bT);        // jtAcc + jt == sT + eT exactly
    const float sg = sT < 0.0f ? -1.0f : 1.0f;
    if ((fabsf(sT) - mh) + (sg * eT - ml) > 0.0f) {                                        // beyond the friction cone: clamp to +-(mh + ml)
        jtS = sg * (mh + ml);
        dtg = (sg * mh - jtO) + sg * ml;
    } else { jtS = sT; dtg = jtI; }
}

// cpArbiterApplyCachedImpulse
__device__ __forceinline__ void contact_warm(CBody &A, CBody &B, float nx, float ny, float jn, float jt, float dt_coef) {
    jn *= dt_coef; jt *= dt_coef;
    float jx = nx * jn - ny * jt, jy = nx * jt + ny * jn;
    A.vx -= jx * A.minv; A.vy -= jy * A.minv; A.w -= A.iinv * A.r * jt;
    B.vx += jx * B.minv; B.vy += jy * B.minv; B.w -= B.iinv * B.r * jt;
}
// cpArbiterApplyImpulse (SURVEY A-5)
__device__ __forceinline__ void contact_apply(CBody &A, CBody &B, float nx, float ny, float bias, float bounce, float mu, float &jnAcc, float &jtAcc, float &jBias) {
    float nMass = 1.0f / (A.minv + B.minv);
    float tMass = 1.0f / (A.minv + B.minv + A.iinv * A.r * A.r + B.iinv * B.r * B.r);
    float dvx = B.vx - A.vx, dvy = B.vy - A.vy;
    float vbn = (B.vbx - A.vbx) * nx + (B.vby - A.vby) * ny;
    float vrn = dvx * nx + dvy * ny;
    float vrt = (dvy * nx - dvx * ny) - B.r * B.w - A.r * A.w;
    float jbn = (bias - vbn) * nMass, jbnOld = jBias;
    jBias = fmaxf(jbnOld + jbn, 0.0f);
    float jnS, jtS, dn, dtg;
    contact_accumulate(jnAcc, jtAcc, -(bounce + vrn) * nMass, -vrt * tMass, mu, jnS, jtS, dn, dtg);
    jnAcc = jnS; jtAcc = jtS;
    float bj = jBias - jbnOld;
    A.vbx -= nx * bj * A.minv; A.vby -= ny * bj * A.minv;
    B.vbx += nx * bj * B.minv; B.vby += ny * bj * B.minv;
    float jx = nx * dn - ny * dtg, jy = nx * dtg + ny * dn;
    A.vx -= jx * A.minv; A.vy -= jy * A.minv; A.w -= A.iinv * A.r * dtg;
    B.vx += jx * B.minv; B.vy += jy * B.minv; B.w -= B.iinv * B.r * dtg;
}

// ------------------------------------------------------------------------------------------------------------------
// cross contacts (global memory)
// ------------------------------------------------------------------------------------------------------------------
struct GRef { float px, py, e, u; CBody b; int kind; };   // kind 0 cell, 1 dead, 2 wall
__device__ __forceinline__ void gload(const SapioWorld &w, uint32_t id, int NC, int ND, GRef &g) {
    if ((int)id < NC) {
        g.kind = 0;
        float2 p = reinterpret_cast<const float2 *>(w.c_pos)[id], v = reinterpret_cast<const float2 *>(w.c_vel)[id], vb = reinterpret_cast<const float2 *>(w.c_vbias)[id];
        float m = w.c_mass[id]; g.b.r = w.c_size[id] * 0.5f;
        g.px = p.x; g.py = p.y; g.e = w.c_elast[id]; g.u = w.c_fric[id];
        g.b.vx = v.x; g.b.vy = v.y; g.b.w = w.c_angvel[id]; g.b.vbx = vb.x; g.b.vby = vb.y;
        g.b.minv = 1.0f / m; g.b.iinv = 1.0f / (0.5f * m * g.b.r * g.b.r);
    } else if ((int)id < NC + ND) {
        g.kind = 1;
        int d = (int)id - NC;
        float2 p = reinterpret_cast<const float2 *>(w.d_pos)[d], v = reinterpret_cast<const float2 *>(w.d_vel)[d], vb = reinterpret_cast<const float2 *>(w.d_vbias)[d];
        g.b.r = w.d_size[d] * 0.5f;
        g.px = p.x; g.py = p.y; g.e = w.d_elast[d]; g.u = w.d_fric[d];
        g.b.vx = v.x; g.b.vy = v.y; g.b.w = w.d_angvel[d]; g.b.vbx = vb.x; g.b.vby = vb.y;
        g.b.minv = 1.0f / w.d_mass[d]; g.b.iinv = 1.0f / (0.5f * w.d_mass0[d] * g.b.r * g.b.r);
    } else {
        g.kind = 2; g.px = g.py = 0.f; g.e = w.p.wall_elast; g.u = w.p.wall_fric;
        g.b = CBody{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, w.p.frame_width};
    }
}
__device__ __forceinline__ void gstore(const SapioWorld &w, uint32_t id, int NC, const GRef &g) {
    if (g.kind == 0) {
        reinterpret_cast<float2 *>(w.c_vel)[id] = make_float2(g.b.vx, g.b.vy); w.c_angvel[id] = g.b.w;
        reinterpret_cast<float2 *>(w.c_vbias)[id] = make_float2(g.b.vbx, g.b.vby);
    } else if (g.kind == 1) {
        int d = (int)id - NC;
        reinterpret_cast<float2 *>(w.d_vel)[d] = make_float2(g.b.vx, g.b.vy); w.d_angvel[d] = g.b.w;
        reinterpret_cast<float2 *>(w.d_vbias)[d] = make_float2(g.b.vbx, g.b.vby);
    }
}
__device__ __forceinline__ bool body_fresh(const SapioWorld &w, uint32_t id, int NC, int ND, int32_t tick) {
    if ((int)id < NC) return w.org_birth_tick[id >> 6] == tick;
    if ((int)id < NC + ND) return w.d_state[id - NC] == 2;
    return false;
}


// ---- connected components of the cross-contact graph (components.cuh): node of a body, and whether its component is one of the
// oversize ones that the cooperative kernel of coupled.cuh solves (cn_tier < 0)
__device__ __forceinline__ int comp_node(uint32_t body, int NC, int NB, int n_org) {          // -1: wall
    if ((int)body < NC) return (int)(body >> 6);
    if ((int)body < NB) return n_org + ((int)body - NC);
    return -1;
}
__device__ __forceinline__ bool comp_is_big(const Workspace &W, int node) {
    int v = node;
    for (;;) { const int p = W.cn_parent[v]; if (p == v) break; v = p; }
    return W.cn_tier[v] < 0;
}
