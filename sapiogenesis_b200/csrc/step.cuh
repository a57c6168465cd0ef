// step.cuh -- pymunk.Space.step(dt) (physics.py:432) on the GPU, for the subset the reference uses:
// circle bodies, four static wall segments, pin joints with both anchors at the body centres.
//
// Stage map (canonical order: oracle/space_step.c header, DESIGN.md "Space.step"):
//   k_integrate        position integration of every body + broadphase key                         (bodies.cuh)
//   radix sort         bodies by grid cell (spatial hash over a dense uniform grid)
//   k_cell_bounds      cell -> [start,end) in the sorted body list
//   k_narrow_find      per body (cell-sorted order): cross partners in the 3x3 neighbourhood (exact double predicate), parked in a scratch run
//   exclusive scans    counts -> offsets: deterministic, globally sorted pair list without sorting pairs
//   k_narrow_fill      per body with partners: its sorted adjacency and its (a<b) contact keys at the scanned offsets
//   k_cross_links      per body: contact index of every adjacency entry + predecessor links (Gauss-Seidel order)
//   k_cross_prestep    per cross contact: warm-start match against the previous step's list, normal/bias/bounce, coupling flags
//   k_classify         organisms -> coupled list, or one of four size classes (16/32/48/64 cell rows)
//   k_org_solve<CM>    one warp per organism without cross contacts: intra contacts, sensor pins, all-pairs pins,
//                      warm start and all iterations with velocities and impulses resident in shared memory
//   k_coupled          cooperative kernel for organisms / dead cells that share a cross contact: dependency levels of
//                      the cross contacts (== exact sequential order evaluated in parallel), then the same arithmetic
//                      through global memory, cross levels interleaved with the per-organism pass, per iteration
//   k_step_finish      publishes the new arbiter list / adjacency, clears "fresh" marks
#pragma once
#include <cooperative_groups.h>
#include <cub/cub.cuh>
#include "common.cuh"
#include "bodies.cuh"
#include "prof.cuh"
namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------------------------
struct Workspace {
    int *pop;                       // population reduction scratch
    int *flags;                     // FL_*
    uint32_t *gkey, *gval, *skey, *sval;   // broadphase keys/values, unsorted and sorted [NB]
    void *cub_tmp; size_t cub_bytes;
    int *cell_start, *cell_end;     // [grid cells]
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
    int *coupled_list;              // [n_org]
    int *sorted_orgs;               // [n_org] uncoupled organisms sorted by cell rows
    int *size_hist;                 // [MAXC + 1] inside the flags block: histogram, then scatter cursors
    float *ic_bounce, *ic_jbias;    // [n_org * ICAP] (coupled path only)
    float4 *pair_cst;               // [4 classes][ORG_GC_WARPS][NPAIR] pin constants of the organism a warp is solving (GC classes)
    char *cp_state; int cp_cap;     // parked solver state of the first cp_cap coupled organisms (sizeof(OrgS<MAXC>) each)
};
enum { FL_NX = 0, FL_NADJ = 1, FL_OVERFLOW = 2, FL_MAXLEVEL = 3, FL_NCOUPLED = 5, FL_NIC = 6, FL_CLS0 = 8 /* ..16: class starts + end */,
       FL_WORK0 = 20 /* ..27 */, FL_TMPTOP = 28, FL_NHIT = 29, FL_HIST = 32 /* ..96 */, FL_CHIST = 100 /* ..164: coupled organisms by rows */,
       FL_CCLS0 = 168 /* ..176: class starts + end in coupled_list */, FL_CHG0 = 180 /* ..182 */, FL_CWORK0 = 200 /* ..203: two alternating pairs of work counters of k_coupled */, FL_BYTES = 1024 };

#define COUPLED_STATE_BYTES 49152      /* >= sizeof(OrgS<MAXC>), checked where the struct is known */
// sort only the bits a grid key can have; the key of an empty slot (all ones) still sorts behind every cell: its low bits are all ones
static inline int grid_key_bits(int ncell) { int b = 1; while (b < 32 && ((1ll << b) - 1) < (long long)ncell) b++; return b; }

static size_t carve_workspace(const SapioParams &p, char *base, Workspace *W) {
    size_t off = 0;
    auto take = [&](size_t bytes) -> char * { char *r = base ? base + off : nullptr; off += (bytes + 255) & ~size_t(255); return r; };
    const size_t NB = (size_t)p.n_org * MAXC + p.n_dead, XC = (size_t)p.x_cap, NG = (size_t)p.grid_nx * p.grid_ny;
    W->pop = (int *)take(256);
    W->flags = (int *)take(FL_BYTES);
    W->size_hist = W->flags ? W->flags + FL_HIST : nullptr;
    W->gkey = (uint32_t *)take(NB * 4); W->gval = (uint32_t *)take(NB * 4);
    W->skey = (uint32_t *)take(NB * 4); W->sval = (uint32_t *)take(NB * 4);
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)NB);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int *)nullptr, (int *)nullptr, (int)NB + 1);
    W->cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
    W->cub_tmp = take(W->cub_bytes);
    W->cell_start = (int *)take(NG * 4); W->cell_end = (int *)take(NG * 4);
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
    W->coupled_list = (int *)take((size_t)p.n_org * 4);
    W->sorted_orgs = (int *)take((size_t)p.n_org * 4);
    W->ic_bounce = (float *)take((size_t)p.n_org * SAPIO_ICAP * 4);
    W->ic_jbias = (float *)take((size_t)p.n_org * SAPIO_ICAP * 4);
    W->pair_cst = (float4 *)take((size_t)4 * (148 * 32) * SAPIO_NPAIR * sizeof(float4));
    W->cp_cap = p.n_org < 65536 ? p.n_org : 65536;
    W->cp_state = take((size_t)W->cp_cap * COUPLED_STATE_BYTES);
    return off;
}

// ------------------------------------------------------------------------------------------------------------------
// broadphase
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cell_bounds(const uint32_t *__restrict__ skey, int n, int ncell, int *__restrict__ cell_start, int *__restrict__ cell_end) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t k = skey[i];
        if (k >= (uint32_t)ncell) continue;
        if (i == 0 || skey[i - 1] != k) cell_start[k] = i;
        if (i == n - 1 || skey[i + 1] != k) cell_end[k] = i + 1;
    }
}

__device__ __forceinline__ void body_circle(const SapioWorld &w, int b, int NC, float &x, float &y, float &r) {
    if (b < NC) { float2 p = reinterpret_cast<const float2 *>(w.c_pos)[b]; x = p.x; y = p.y; r = w.c_size[b] * 0.5f; }
    else { int d = b - NC; float2 p = reinterpret_cast<const float2 *>(w.d_pos)[d]; x = p.x; y = p.y; r = w.d_size[d] * 0.5f; }
}

#define NARROW_LOCAL 48
// Pass 1, one thread per entry of the cell-sorted body list (neighbouring threads walk the same grid cells). Partners = solid
// bodies that are not living cells of the same organism (those pairs belong to the organism kernel), found in the 3x3 grid
// cells around the body; dead cells also test the four walls. A body with partners parks them (ascending id, walls last) in a
// scratch run and registers itself in the hit list; almost every body has none.
__global__ void __launch_bounds__(128) k_narrow_find(WORLD_ARG, Workspace W) {
    const SapioParams &p = w.p;
    const int NC = p.n_org * MAXC, NB = NC + p.n_dead, WALL0 = NB;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < NB; idx += gridDim.x * blockDim.x) {
        const uint32_t key = W.skey[idx];
        if (key == 0xFFFFFFFFu) continue;                    // empty slots sort to the end
        const int b = (int)W.sval[idx];
        float mx, my, mr;
        body_circle(w, b, NC, mx, my, mr);
        const int my_org = b < NC ? (b >> 6) : -1;
        const int cx = (int)(key % (uint32_t)p.grid_nx), cy = (int)(key / (uint32_t)p.grid_nx);
        uint32_t loc[NARROW_LOCAL];
        int n = 0, up = 0;
        bool over = false;
        for (int yy = cy - 1; yy <= cy + 1; yy++) {
            if (yy < 0 || yy >= p.grid_ny) continue;
            for (int xx = cx - 1; xx <= cx + 1; xx++) {
                if (xx < 0 || xx >= p.grid_nx) continue;
                int c = yy * p.grid_nx + xx;
                int s = W.cell_start[c], e = W.cell_end[c];
                for (int i = s; i < e; i++) {
                    int j = (int)W.sval[i];
                    if (j == b) continue;
                    if (my_org >= 0 && j < NC && (j >> 6) == my_org) continue;
                    float ox, oy, orad;
                    body_circle(w, j, NC, ox, oy, orad);
                    double d2;
                    if (!circles_overlap(mx, my, mr, ox, oy, orad, &d2)) continue;
                    if (n < NARROW_LOCAL) { loc[n] = (uint32_t)j; n++; up += (j > b); } else over = true;
                }
            }
        }
        if (b >= NC) {
            for (int wi = 0; wi < SAPIO_NWALL; wi++) {
                WallHit h;
                if (!circle_wall(p, wi, mx, my, mr, &h)) continue;
                if (n < NARROW_LOCAL) { loc[n] = (uint32_t)(WALL0 + wi); n++; up++; } else over = true;
            }
        }
        if (over) atomicOr(&W.flags[FL_OVERFLOW], 1);
        if (n == 0) continue;
        W.cnt_all[b] = n; W.cnt_up[b] = up;
        for (int i = 1; i < n; i++) { uint32_t v = loc[i]; int j = i - 1; while (j >= 0 && loc[j] > v) { loc[j + 1] = loc[j]; j--; } loc[j + 1] = v; }
        const int start = atomicAdd(&W.flags[FL_TMPTOP], n), h = atomicAdd(&W.flags[FL_NHIT], 1);
        if (start + n > 2 * p.x_cap || h >= 2 * p.x_cap) { atomicOr(&W.flags[FL_OVERFLOW], 1); continue; }
        W.hit_list[h] = make_int2(b, start);
        for (int i = 0; i < n; i++) W.adj_tmp[start + i] = loc[i];
    }
}
// Pass 2, after the scans of the per-body counts: every registered body copies its run to its place in the adjacency and emits
// the contact keys (partners with a larger id) -- both lists come out globally sorted without sorting pairs.
__global__ void __launch_bounds__(128) k_narrow_fill(WORLD_ARG, Workspace W) {
    const int nh = min(W.flags[FL_NHIT], 2 * w.p.x_cap), xcap = w.p.x_cap;
    for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < nh; h += gridDim.x * blockDim.x) {
        const int2 e = W.hit_list[h];
        const int b = e.x, n = W.cnt_all[b];
        const int oa = W.off_all[b];
        int ou = W.off_up[b];
        for (int i = 0; i < n; i++) {
            const uint32_t v = W.adj_tmp[e.y + i];
            if (oa + i < 2 * xcap) W.adj_other[oa + i] = v;
            if (v > (uint32_t)b) { if (ou < xcap) W.nx_key[ou] = ((uint64_t)(uint32_t)b << 32) | v; ou++; }
        }
    }
}

// scans done: publish totals, detect capacity overflow
__global__ void k_narrow_totals(WORLD_ARG, Workspace W) {
    const int NB = w.p.n_org * MAXC + w.p.n_dead;
    int nx = W.off_up[NB], na = W.off_all[NB];
    if (nx > w.p.x_cap || na > 2 * w.p.x_cap) { W.flags[FL_OVERFLOW] = 1; nx = min(nx, w.p.x_cap); na = min(na, 2 * w.p.x_cap); }
    W.flags[FL_NX] = nx; W.flags[FL_NADJ] = na;
}

// ------------------------------------------------------------------------------------------------------------------
// contact arithmetic (cpArbiter.c) specialised for circles: every contact offset is r1 = ra*n, r2 = -rb*n (a wall
// counts as rb = frame radius on a static body), so r x n = 0 exactly. Writing that out removes the catastrophic
// cancellation the general r x j form has in fp32:
//   k(n) = ma + mb            k(t) = ma + mb + ia*ra^2 + ib*rb^2
//   v_rel.n = (vb - va).n     v_rel.t = (vb - va).t - rb*wb - ra*wa
//   torque of impulse (jn, jt): a: -ia*ra*jt   b: -ib*rb*jt      bias impulses (along n) never touch w_bias
// ------------------------------------------------------------------------------------------------------------------
struct StepK { double bc_contact, bc_joint, dt; float dt_coef_unused; };

struct CBody { float vx, vy, w, vbx, vby, minv, iinv, r; };

__device__ __forceinline__ float contact_bias(double d, float ra, float rb, const SapioParams &p, const StepK &K) {
    double dist = d - (double)ra - (double)rb;                            // ((r2 - r1) + (pb - pa)) . n
    return (float)(-K.bc_contact * fmin(0.0, dist + (double)p.collision_slop) / K.dt);
}
__device__ __forceinline__ float contact_vrn(const CBody &A, const CBody &B, float nx, float ny) { return (B.vx - A.vx) * nx + (B.vy - A.vy) * ny; }

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
    // accumulate and clamp in double, apply the difference (see OrgCtx::contacts_pass)
    const double jnOld = (double)jnAcc, jnNew = fmax(jnOld + (double)(-(bounce + vrn) * nMass), 0.0);
    const double jtMax = (double)mu * jnNew;
    const double jtOld = (double)jtAcc, jtNew = fmin(fmax(jtOld + (double)(-vrt * tMass), -jtMax), jtMax);
    jnAcc = (float)jnNew; jtAcc = (float)jtNew;
    float bj = jBias - jbnOld;
    A.vbx -= nx * bj * A.minv; A.vby -= ny * bj * A.minv;
    B.vbx += nx * bj * B.minv; B.vby += ny * bj * B.minv;
    float dn = (float)(jnNew - jnOld), dtg = (float)(jtNew - jtOld);
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

// per body: predecessor links of its contacts in canonical order (its adjacency order); contact index by construction
__global__ void __launch_bounds__(128) k_cross_links(WORLD_ARG, Workspace W) {
    const int NC = w.p.n_org * MAXC, NB = NC + w.p.n_dead;
    const int nx = W.flags[FL_NX];
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < NB; b += gridDim.x * blockDim.x) {
        int s = W.off_all[b], e = W.off_all[b + 1];
        if (e > 2 * w.p.x_cap) e = 2 * w.p.x_cap;
        int prev = -1, nup = 0;
        for (int q = s; q < e; q++) {
            uint32_t y = W.adj_other[q];
            int c;
            if (y > (uint32_t)b) c = W.off_up[b] + nup++;
            else {   // contact (y, b): find b among y's larger partners
                int lo = W.off_up[y], hi = min(W.off_up[y + 1], nx);
                uint64_t key = ((uint64_t)y << 32) | (uint32_t)b;
                while (lo < hi) { int mid = (lo + hi) >> 1; if (W.nx_key[mid] < key) lo = mid + 1; else hi = mid; }
                c = lo;
            }
            W.adj_c[q] = c < nx ? c : -1;
            if (c < nx) { if (y > (uint32_t)b) W.nx_prev_a[c] = prev; else W.nx_prev_b[c] = prev; }
            prev = c;
        }
    }
}

// per new cross contact: persistent-arbiter match (jnAcc/jtAcc of the previous step), normal and bias (double, once),
// bounce from pre-warm-start velocities (cpArbiterPreStep), coupling marks
__global__ void __launch_bounds__(128) k_cross_prestep(WORLD_ARG, Workspace W, StepK K) {
    const SapioParams &p = w.p;
    const int NC = p.n_org * MAXC, ND = p.n_dead, WALL0 = NC + ND;
    const int nx = W.flags[FL_NX];
    const int32_t tick = (int32_t)w.g_tick[0];
    const int pn = w.g_x_n[0];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nx; c += gridDim.x * blockDim.x) {
        uint64_t key = W.nx_key[c];
        uint32_t a = (uint32_t)(key >> 32), b = (uint32_t)key;
        GRef A, B; gload(w, a, NC, ND, A); gload(w, b, NC, ND, B);
        float nxn, nyn; double d;
        if (B.kind == 2) { WallHit h; circle_wall(p, (int)b - WALL0, A.px, A.py, A.b.r, &h); nxn = h.nx; nyn = h.ny; d = h.d; }
        else {
            double dx = (double)B.px - (double)A.px, dy = (double)B.py - (double)A.py;
            d = sqrt(dx * dx + dy * dy);
            if (d != 0.0) { nxn = (float)(dx / d); nyn = (float)(dy / d); } else { nxn = 1.f; nyn = 0.f; }
        }
        W.nx_n[c] = make_float2(nxn, nyn);
        W.nx_bias[c] = contact_bias(d, A.b.r, B.b.r, p, K);
        W.nx_bounce[c] = contact_vrn(A.b, B.b, nxn, nyn) * (A.e * B.e);
        W.nx_jbias[c] = 0.f;
        float jn = 0.f, jt = 0.f;
        if (pn > 0 && !body_fresh(w, a, NC, ND, tick) && !body_fresh(w, b, NC, ND, tick)) {
            int lo = 0, hi = pn;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (w.x_key[mid] < key) lo = mid + 1; else hi = mid; }
            if (lo < pn && w.x_key[lo] == key) { jn = w.x_jn[lo]; jt = w.x_jt[lo]; }
        }
        W.nx_jn[c] = jn; W.nx_jt[c] = jt;
        W.nx_level[c] = 1;
        if (A.kind == 0) W.org_coupled[a >> 6] = 1;
        if (B.kind == 0) W.org_coupled[b >> 6] = 1;
    }
}

#include "orgsolve.cuh"

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
#define COUPLED_PACK_FROM 2400            /* coupled organisms (last tick) from which the packed form is launched: ~4 per warp of the simple one */
static int g_coupled_pack_from = COUPLED_PACK_FROM;       // sapio_set_coupled_pack_from
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

// publish: the new arbiter list becomes w.x_*, the adjacency w.xa_*; created-this-tick marks are cleared
__global__ void __launch_bounds__(256) k_step_finish(WORLD_ARG, Workspace W, int *mail) {
    const int NC = w.p.n_org * MAXC, ND = w.p.n_dead, NB = NC + ND;
    const int nx = W.flags[FL_NX], na = W.flags[FL_NADJ];
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    for (int c = gtid; c < nx; c += gthreads) { w.x_key[c] = W.nx_key[c]; w.x_jn[c] = W.nx_jn[c]; w.x_jt[c] = W.nx_jt[c]; }
    for (int q = gtid; q < na; q += gthreads) w.xa_other[q] = W.adj_other[q];
    for (int b = gtid; b <= NB; b += gthreads) w.xa_off[b] = min(W.off_all[b], na);
    for (int d = gtid; d < ND; d += gthreads) if (w.d_state[d] == 2) w.d_state[d] = 1;
    if (gtid == 0) {
        w.g_x_n[0] = nx; w.g_stepped[0] = 1;
        if (mail) mail[0] = W.flags[FL_NCOUPLED];
        w.g_stats[SAPIO_STAT_XCONTACTS] = nx; w.g_stats[SAPIO_STAT_ICONTACTS] = W.flags[FL_NIC];
        w.g_stats[SAPIO_STAT_LEVELS] = W.flags[FL_MAXLEVEL]; w.g_stats[SAPIO_STAT_COUPLED] = W.flags[FL_NCOUPLED];
        if (W.flags[FL_OVERFLOW]) w.g_stats[SAPIO_STAT_OVERFLOW] += 1;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host-side launch sequence
// ------------------------------------------------------------------------------------------------------------------
struct StepPlan { bool ready; int sms; int coop_blocks, simple_blocks; size_t simple_smem; int *h_mail, *d_mail; int org_blocks[ORG_CLASSES]; size_t org_smem[ORG_CLASSES]; size_t coop_smem; cudaStream_t side[ORG_CLASSES]; cudaEvent_t fork, join[ORG_CLASSES]; };
static StepPlan g_plan = {};

template <int CM>
static cudaError_t plan_org(int cls) {
    size_t smem = (size_t)OrgCfg<CM>::WARP_BYTES * OrgCfg<CM>::WARPS;
    cudaError_t e = cudaFuncSetAttribute(k_org_solve<CM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_org_solve<CM>, OrgCfg<CM>::WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    g_plan.org_smem[cls] = smem;
    g_plan.org_blocks[cls] = (per_sm > 0 ? per_sm : 1) * g_plan.sms;     // persistent: every SM full, work pulled dynamically
    return cudaSuccess;
}

static int launch_space_step(const SapioWorld &w, const Workspace &W, cudaStream_t st, char *err, size_t errlen) {
#define STEP_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { snprintf(err, errlen, #expr ": %s", cudaGetErrorString(e_)); return SAPIO_E_CUDA; } } while (0)
    const SapioParams &p = w.p;
    const int NC = p.n_org * MAXC, NB = NC + p.n_dead, NG = p.grid_nx * p.grid_ny;
    if (!g_plan.ready) {
        int dev = 0;
        STEP_TRY(cudaGetDevice(&dev));
        STEP_TRY(cudaDeviceGetAttribute(&g_plan.sms, cudaDevAttrMultiProcessorCount, dev));
        STEP_TRY(plan_org<8>(0)); STEP_TRY(plan_org<16>(1)); STEP_TRY(plan_org<24>(2)); STEP_TRY(plan_org<32>(3));
        STEP_TRY(plan_org<40>(4)); STEP_TRY(plan_org<48>(5)); STEP_TRY(plan_org<56>(6)); STEP_TRY(plan_org<64>(7));
        STEP_TRY(cudaEventCreateWithFlags(&g_plan.fork, cudaEventDisableTiming));
        for (int c = 0; c < ORG_CLASSES; c++) {
            STEP_TRY(cudaStreamCreateWithFlags(&g_plan.side[c], cudaStreamNonBlocking));
            STEP_TRY(cudaEventCreateWithFlags(&g_plan.join[c], cudaEventDisableTiming));
        }
        g_plan.simple_smem = sizeof(OrgS<MAXC>) * SIMPLE_WARPS;
        STEP_TRY(cudaFuncSetAttribute(k_coupled_simple, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_plan.simple_smem));
        { int ps = 0; STEP_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ps, k_coupled_simple, SIMPLE_WARPS * 32, g_plan.simple_smem)); g_plan.simple_blocks = ps * g_plan.sms; }
        if (g_plan.simple_blocks < 1) { snprintf(err, errlen, "k_coupled_simple does not fit on an SM"); return SAPIO_E_CUDA; }
        // mailbox: the device leaves the number of coupled organisms of the tick here (mapped host memory); the host reads it, one tick
        // stale and without waiting, to pick the form of the coupled pass
        STEP_TRY(cudaHostAlloc((void **)&g_plan.h_mail, 64, cudaHostAllocMapped));
        STEP_TRY(cudaHostGetDevicePointer((void **)&g_plan.d_mail, g_plan.h_mail, 0));
        g_plan.h_mail[0] = 0;
        g_plan.coop_smem = CP_SLAB * COUPLED_WARPS;
        STEP_TRY(cudaFuncSetAttribute(k_coupled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_plan.coop_smem));
        int per_sm = 0;
        STEP_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_coupled, COUPLED_WARPS * 32, g_plan.coop_smem));
        g_plan.coop_blocks = per_sm * g_plan.sms;
        if (g_plan.coop_blocks < 1) { snprintf(err, errlen, "k_coupled does not fit on an SM"); return SAPIO_E_CUDA; }
        g_plan.ready = true;
    }
    const int sms = g_plan.sms;
    auto sgrid = [&](long items, int threads) { long want = (items + threads - 1) / threads, wave = (long)sms * 8; return (int)(want < 1 ? 1 : (want > wave ? wave : want)); };
    StepK K;
    K.dt = (double)p.dt_phys;
    K.bc_contact = 1.0 - pow((double)p.collision_bias, K.dt);        // cpSpaceStep: biasCoef = 1 - collisionBias^dt
    K.bc_joint = 1.0 - pow((double)p.joint_error_bias, K.dt);        // bias_coef(errorBias, dt)
    K.dt_coef_unused = 0.f;

    prof_mark(st, PT_BROADPHASE);
    STEP_TRY(cudaMemsetAsync(W.flags, 0, FL_BYTES, st));
    STEP_TRY(cudaMemsetAsync(W.org_coupled, 0, (size_t)p.n_org, st));
    k_integrate<true><<<sgrid(NB, 256), 256, 0, st>>>(w, W.gkey, W.gval);
    size_t tmp = W.cub_bytes;
    STEP_TRY(cub::DeviceRadixSort::SortPairs(W.cub_tmp, tmp, W.gkey, W.skey, W.gval, W.sval, NB, 0, grid_key_bits(NG), st));
    STEP_TRY(cudaMemsetAsync(W.cell_start, 0, (size_t)NG * 4, st));
    STEP_TRY(cudaMemsetAsync(W.cell_end, 0, (size_t)NG * 4, st));
    k_cell_bounds<<<sgrid(NB, 256), 256, 0, st>>>(W.skey, NB, NG, W.cell_start, W.cell_end);
    prof_mark(st, PT_NARROW);
    STEP_TRY(cudaMemsetAsync(W.cnt_up, 0, (size_t)(NB + 1) * 4, st));
    STEP_TRY(cudaMemsetAsync(W.cnt_all, 0, (size_t)(NB + 1) * 4, st));
    k_narrow_find<<<sgrid(NB, 128), 128, 0, st>>>(w, W);
    tmp = W.cub_bytes;
    STEP_TRY(cub::DeviceScan::ExclusiveSum(W.cub_tmp, tmp, W.cnt_up, W.off_up, NB + 1, st));
    tmp = W.cub_bytes;
    STEP_TRY(cub::DeviceScan::ExclusiveSum(W.cub_tmp, tmp, W.cnt_all, W.off_all, NB + 1, st));
    k_narrow_totals<<<1, 1, 0, st>>>(w, W);
    k_narrow_fill<<<sgrid(2 * (long)p.x_cap, 128), 128, 0, st>>>(w, W);
    STEP_TRY(cudaMemsetAsync(W.nx_prev_a, 0xFF, (size_t)p.x_cap * 4, st));
    STEP_TRY(cudaMemsetAsync(W.nx_prev_b, 0xFF, (size_t)p.x_cap * 4, st));
    k_cross_links<<<sgrid(NB, 128), 128, 0, st>>>(w, W);
    k_cross_prestep<<<sgrid(p.x_cap, 128), 128, 0, st>>>(w, W, K);
    k_classify_count<<<sgrid(p.n_org, 256), 256, 0, st>>>(w, W);
    k_classify_scan<<<1, 32, 0, st>>>(W);
    k_classify_scatter<<<sgrid(p.n_org, 256), 256, 0, st>>>(w, W);
    LAUNCHES(11);
    prof_mark(st, PT_ORG_SOLVE);
    {   // persistent CTAs per size class, largest class first; the classes run concurrently on side streams so that the tail of
        // one class is filled by the next (organisms of different classes are independent)
        STEP_TRY(cudaEventRecord(g_plan.fork, st));
#define ORG_LAUNCH(CM, cls) do { \
            cudaStream_t ss = g_plan.side[cls]; \
            STEP_TRY(cudaStreamWaitEvent(ss, g_plan.fork, 0)); \
            k_org_solve<CM><<<g_plan.org_blocks[cls], OrgCfg<CM>::WARPS * 32, g_plan.org_smem[cls], ss>>>(w, W, K, cls); \
            STEP_TRY(cudaEventRecord(g_plan.join[cls], ss)); \
            STEP_TRY(cudaStreamWaitEvent(st, g_plan.join[cls], 0)); } while (0)
        ORG_LAUNCH(64, 7); ORG_LAUNCH(56, 6); ORG_LAUNCH(48, 5); ORG_LAUNCH(40, 4);
        ORG_LAUNCH(32, 3); ORG_LAUNCH(24, 2); ORG_LAUNCH(16, 1); ORG_LAUNCH(8, 0);
#undef ORG_LAUNCH
    }
    LAUNCHES(8);
    prof_mark(st, PT_COUPLED);
    {
        SapioWorld wc = w; Workspace Wc = W; StepK Kc = K;
        void *args[] = { (void *)&wc, (void *)&Wc, (void *)&Kc };
        const int last = *(volatile int *)g_plan.h_mail;
        if (last > g_coupled_pack_from) STEP_TRY(cudaLaunchCooperativeKernel((void *)k_coupled, dim3(g_plan.coop_blocks), dim3(COUPLED_WARPS * 32), args, g_plan.coop_smem, st));
        else STEP_TRY(cudaLaunchCooperativeKernel((void *)k_coupled_simple, dim3(g_plan.simple_blocks), dim3(SIMPLE_WARPS * 32), args, g_plan.simple_smem, st));
    }
    prof_mark(st, PT_STEP_FINISH);
    k_step_finish<<<sgrid(NB + 1, 256), 256, 0, st>>>(w, W, g_plan.d_mail);
    LAUNCHES(2);
    prof_mark(st, PT_END);
    STEP_TRY(cudaGetLastError());
    return 0;
#undef STEP_TRY
}
