// step.cuh -- pymunk.Space.step(dt) (physics.py:432) on the GPU, for the subset the reference uses:
// circle bodies, four static wall segments, pin joints with both anchors at the body centres.
//
// Stage map (canonical order: oracle/space_step.c header, DESIGN.md "Space.step"):
//   k_integrate        position integration of every body + broadphase key                         (bodies.cuh)
//   radix sort         bodies by grid cell (spatial hash over a dense uniform grid)
//   k_cell_bounds      cell -> [start,end) in the sorted body list
//   k_narrow<false>    per body: cross partners in the 3x3 neighbourhood (exact double predicate), counts
//   exclusive scans    counts -> offsets: deterministic, globally sorted pair list without sorting pairs
//   k_narrow<true>     per body: writes its sorted adjacency and its (a<b) contact keys
//   k_cross_links      per body: contact index of every adjacency entry + predecessor links (Gauss-Seidel order)
//   k_cross_prestep    per cross contact: warm-start match against the previous step's list, bounce, coupling flags
//   k_org_solve        one warp per organism without cross contacts: intra contacts, sensor pins, all-pairs pins,
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
    float *nx_jn, *nx_jt, *nx_bounce, *nx_jbias;
    int *nx_level, *nx_prev_a, *nx_prev_b;
    uint32_t *adj_other;            // adjacency: partner body id per entry [2*x_cap]
    uint8_t *org_coupled;           // [n_org]
    int *coupled_list;              // [n_org]
    float *ic_bounce, *ic_jbias;    // [n_org * ICAP] (coupled path only)
};
enum { FL_NX = 0, FL_NADJ = 1, FL_OVERFLOW = 2, FL_MAXLEVEL = 3, FL_CHANGED = 4, FL_NCOUPLED = 5, FL_NIC = 6 };

static size_t carve_workspace(const SapioParams &p, char *base, Workspace *W) {
    size_t off = 0;
    auto take = [&](size_t bytes) -> char * { char *r = base ? base + off : nullptr; off += (bytes + 255) & ~size_t(255); return r; };
    const size_t NB = (size_t)p.n_org * MAXC + p.n_dead, XC = (size_t)p.x_cap, NG = (size_t)p.grid_nx * p.grid_ny;
    W->pop = (int *)take(256);
    W->flags = (int *)take(256);
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
    W->nx_level = (int *)take(XC * 4); W->nx_prev_a = (int *)take(XC * 4); W->nx_prev_b = (int *)take(XC * 4);
    W->adj_other = (uint32_t *)take(2 * XC * 4);
    W->org_coupled = (uint8_t *)take((size_t)p.n_org);
    W->coupled_list = (int *)take((size_t)p.n_org * 4);
    W->ic_bounce = (float *)take((size_t)p.n_org * SAPIO_ICAP * 4);
    W->ic_jbias = (float *)take((size_t)p.n_org * SAPIO_ICAP * 4);
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
// One thread per body slot. Partners = solid bodies that are not living cells of the same organism (those pairs belong
// to the organism kernel), found in the 3x3 grid cells around the body; dead cells also test the four walls.
// FILL = false: counts. FILL = true: writes the adjacency (all partners, ascending id, walls last) and the contact keys
// (partners with a larger id) at the scanned offsets -- both lists come out globally sorted without sorting pairs.
template <bool FILL>
__global__ void __launch_bounds__(128) k_narrow(WORLD_ARG, Workspace W) {
    const SapioParams &p = w.p;
    const int NC = p.n_org * MAXC, NB = NC + p.n_dead, WALL0 = NB;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b <= NB; b += gridDim.x * blockDim.x) {
        if (b == NB) { if (!FILL) { W.cnt_up[b] = 0; W.cnt_all[b] = 0; } continue; }
        uint32_t key = W.gkey[b];
        if (key == 0xFFFFFFFFu) { if (!FILL) { W.cnt_up[b] = 0; W.cnt_all[b] = 0; } continue; }
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
                    if (n < NARROW_LOCAL) { if (FILL) loc[n] = (uint32_t)j; n++; up += (j > b); } else over = true;
                }
            }
        }
        if (b >= NC) {
            for (int wi = 0; wi < SAPIO_NWALL; wi++) {
                WallHit h;
                if (!circle_wall(p, wi, mx, my, mr, &h)) continue;
                if (n < NARROW_LOCAL) { if (FILL) loc[n] = (uint32_t)(WALL0 + wi); n++; up++; } else over = true;
            }
        }
        if (over) atomicOr(&W.flags[FL_OVERFLOW], 1);
        if (!FILL) { W.cnt_all[b] = n; W.cnt_up[b] = up; }
        else {
            for (int i = 1; i < n; i++) { uint32_t v = loc[i]; int j = i - 1; while (j >= 0 && loc[j] > v) { loc[j + 1] = loc[j]; j--; } loc[j + 1] = v; }
            int oa = W.off_all[b], ou = W.off_up[b];
            const int xcap = p.x_cap;
            for (int i = 0; i < n; i++) {
                if (oa + i < 2 * xcap) W.adj_other[oa + i] = loc[i];
                if (loc[i] > (uint32_t)b) { if (ou < xcap) W.nx_key[ou] = ((uint64_t)(uint32_t)b << 32) | loc[i]; ou++; }
            }
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
// contact arithmetic shared by the shared-memory (organism) and global-memory (cross) paths: cpArbiter.c
// ------------------------------------------------------------------------------------------------------------------
struct CBody { float vx, vy, w, vbx, vby, wb, minv, iinv; };
struct CGeo { float nx, ny, r1x, r1y, r2x, r2y, nMass, tMass, bias; };

__device__ __forceinline__ float k_scalar(float minv_a, float iinv_a, float minv_b, float iinv_b, float r1x, float r1y, float r2x, float r2y, float nx, float ny) {
    float c1 = r1x * ny - r1y * nx, c2 = r2x * ny - r2y * nx;
    return (minv_a + iinv_a * c1 * c1) + (minv_b + iinv_b * c2 * c2);
}
// geometry of a circle/circle contact (a at pa radius ra, b at pb radius rb); cpCollision.c CircleToCircle + cpArbiterPreStep
__device__ __forceinline__ void geo_circles(float pax, float pay, float ra, float pbx, float pby, float rb, float minv_a, float iinv_a, float minv_b, float iinv_b,
                                            float slop, float bias_coef, float inv_dt, CGeo &g) {
    float dx = pbx - pax, dy = pby - pay;
    float d = sqrtf(dx * dx + dy * dy);
    if (d != 0.f) { float inv = 1.0f / d; g.nx = dx * inv; g.ny = dy * inv; } else { g.nx = 1.f; g.ny = 0.f; }
    g.r1x = g.nx * ra; g.r1y = g.ny * ra; g.r2x = -g.nx * rb; g.r2y = -g.ny * rb;
    g.nMass = 1.0f / k_scalar(minv_a, iinv_a, minv_b, iinv_b, g.r1x, g.r1y, g.r2x, g.r2y, g.nx, g.ny);
    g.tMass = 1.0f / k_scalar(minv_a, iinv_a, minv_b, iinv_b, g.r1x, g.r1y, g.r2x, g.r2y, -g.ny, g.nx);
    float dist = ((g.r2x - g.r1x) + dx) * g.nx + ((g.r2y - g.r1y) + dy) * g.ny;
    g.bias = -bias_coef * fminf(0.0f, dist + slop) * inv_dt;
}
// circle against a wall (static body at the origin); h from circle_wall()
__device__ __forceinline__ void geo_wall(float pax, float pay, float ra, const WallHit &h, float fw, float minv_a, float iinv_a, float slop, float bias_coef, float inv_dt, CGeo &g) {
    g.nx = h.nx; g.ny = h.ny;
    g.r1x = g.nx * ra; g.r1y = g.ny * ra;
    g.r2x = h.cx - g.nx * fw; g.r2y = h.cy - g.ny * fw;
    g.nMass = 1.0f / k_scalar(minv_a, iinv_a, 0.f, 0.f, g.r1x, g.r1y, g.r2x, g.r2y, g.nx, g.ny);
    g.tMass = 1.0f / k_scalar(minv_a, iinv_a, 0.f, 0.f, g.r1x, g.r1y, g.r2x, g.r2y, -g.ny, g.nx);
    float dist = ((g.r2x - g.r1x) + (0.f - pax)) * g.nx + ((g.r2y - g.r1y) + (0.f - pay)) * g.ny;
    g.bias = -bias_coef * fminf(0.0f, dist + slop) * inv_dt;
}
__device__ __forceinline__ float contact_bounce(const CBody &A, const CBody &B, const CGeo &g, float e) {
    float rvx = (B.vx - g.r2y * B.w) - (A.vx - g.r1y * A.w), rvy = (B.vy + g.r2x * B.w) - (A.vy + g.r1x * A.w);
    return (rvx * g.nx + rvy * g.ny) * e;
}
__device__ __forceinline__ void apply_imp(CBody &A, CBody &B, const CGeo &g, float jx, float jy) {
    A.vx -= jx * A.minv; A.vy -= jy * A.minv; A.w -= A.iinv * (g.r1x * jy - g.r1y * jx);
    B.vx += jx * B.minv; B.vy += jy * B.minv; B.w += B.iinv * (g.r2x * jy - g.r2y * jx);
}
// cpArbiterApplyCachedImpulse
__device__ __forceinline__ void contact_warm(CBody &A, CBody &B, const CGeo &g, float jn, float jt, float dt_coef) {
    float jx = (g.nx * jn - g.ny * jt) * dt_coef, jy = (g.nx * jt + g.ny * jn) * dt_coef;
    apply_imp(A, B, g, jx, jy);
}
// cpArbiterApplyImpulse (SURVEY A-5)
__device__ __forceinline__ void contact_apply(CBody &A, CBody &B, const CGeo &g, float bounce, float mu, float &jnAcc, float &jtAcc, float &jBias) {
    float vb1x = A.vbx - g.r1y * A.wb, vb1y = A.vby + g.r1x * A.wb;
    float vb2x = B.vbx - g.r2y * B.wb, vb2y = B.vby + g.r2x * B.wb;
    float rvx = (B.vx - g.r2y * B.w) - (A.vx - g.r1y * A.w), rvy = (B.vy + g.r2x * B.w) - (A.vy + g.r1x * A.w);
    float vbn = (vb2x - vb1x) * g.nx + (vb2y - vb1y) * g.ny;
    float vrn = rvx * g.nx + rvy * g.ny;
    float vrt = -rvx * g.ny + rvy * g.nx;
    float jbn = (g.bias - vbn) * g.nMass, jbnOld = jBias;
    jBias = fmaxf(jbnOld + jbn, 0.0f);
    float jn = -(bounce + vrn) * g.nMass, jnOld = jnAcc;
    jnAcc = fmaxf(jnOld + jn, 0.0f);
    float jtMax = mu * jnAcc;
    float jt = -vrt * g.tMass, jtOld = jtAcc;
    jtAcc = fminf(fmaxf(jtOld + jt, -jtMax), jtMax);
    float bj = jBias - jbnOld;
    float bjx = g.nx * bj, bjy = g.ny * bj;
    A.vbx -= bjx * A.minv; A.vby -= bjy * A.minv; A.wb -= A.iinv * (g.r1x * bjy - g.r1y * bjx);
    B.vbx += bjx * B.minv; B.vby += bjy * B.minv; B.wb += B.iinv * (g.r2x * bjy - g.r2y * bjx);
    float dn = jnAcc - jnOld, dtg = jtAcc - jtOld;
    apply_imp(A, B, g, g.nx * dn - g.ny * dtg, g.nx * dtg + g.ny * dn);
}

__device__ __forceinline__ float bias_coef_of(float bias, float dt) { return 1.0f - powf(bias, dt); }

// ------------------------------------------------------------------------------------------------------------------
// cross contacts (global memory)
// ------------------------------------------------------------------------------------------------------------------
struct GRef { float px, py, r, e, u; CBody b; int kind; };   // kind 0 cell, 1 dead, 2 wall
__device__ __forceinline__ void gload(const SapioWorld &w, uint32_t id, int NC, int ND, GRef &g) {
    if ((int)id < NC) {
        g.kind = 0;
        float2 p = reinterpret_cast<const float2 *>(w.c_pos)[id], v = reinterpret_cast<const float2 *>(w.c_vel)[id], vb = reinterpret_cast<const float2 *>(w.c_vbias)[id];
        float m = w.c_mass[id]; g.r = w.c_size[id] * 0.5f;
        g.px = p.x; g.py = p.y; g.e = w.c_elast[id]; g.u = w.c_fric[id];
        g.b.vx = v.x; g.b.vy = v.y; g.b.w = w.c_angvel[id]; g.b.vbx = vb.x; g.b.vby = vb.y; g.b.wb = w.c_wbias[id];
        g.b.minv = 1.0f / m; g.b.iinv = 1.0f / (0.5f * m * g.r * g.r);
    } else if ((int)id < NC + ND) {
        g.kind = 1;
        int d = (int)id - NC;
        float2 p = reinterpret_cast<const float2 *>(w.d_pos)[d], v = reinterpret_cast<const float2 *>(w.d_vel)[d], vb = reinterpret_cast<const float2 *>(w.d_vbias)[d];
        g.r = w.d_size[d] * 0.5f;
        g.px = p.x; g.py = p.y; g.e = w.d_elast[d]; g.u = w.d_fric[d];
        g.b.vx = v.x; g.b.vy = v.y; g.b.w = w.d_angvel[d]; g.b.vbx = vb.x; g.b.vby = vb.y; g.b.wb = w.d_wbias[d];
        g.b.minv = 1.0f / w.d_mass[d]; g.b.iinv = 1.0f / (0.5f * w.d_mass0[d] * g.r * g.r);
    } else {
        g.kind = 2; g.px = g.py = 0.f; g.r = w.p.frame_width; g.e = w.p.wall_elast; g.u = w.p.wall_fric;
        g.b = CBody{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    }
}
__device__ __forceinline__ void gstore(const SapioWorld &w, uint32_t id, int NC, const GRef &g) {
    if (g.kind == 0) {
        reinterpret_cast<float2 *>(w.c_vel)[id] = make_float2(g.b.vx, g.b.vy); w.c_angvel[id] = g.b.w;
        reinterpret_cast<float2 *>(w.c_vbias)[id] = make_float2(g.b.vbx, g.b.vby); w.c_wbias[id] = g.b.wb;
    } else if (g.kind == 1) {
        int d = (int)id - NC;
        reinterpret_cast<float2 *>(w.d_vel)[d] = make_float2(g.b.vx, g.b.vy); w.d_angvel[d] = g.b.w;
        reinterpret_cast<float2 *>(w.d_vbias)[d] = make_float2(g.b.vbx, g.b.vby); w.d_wbias[d] = g.b.wb;
    }
}
__device__ __forceinline__ void cross_geo(const SapioWorld &w, const GRef &A, const GRef &B, uint32_t idb, int WALL0, float bias_coef, CGeo &g) {
    const float inv_dt = 1.0f / w.p.dt_phys;
    if (B.kind == 2) {
        WallHit h; circle_wall(w.p, (int)idb - WALL0, A.px, A.py, A.r, &h);
        geo_wall(A.px, A.py, A.r, h, w.p.frame_width, A.b.minv, A.b.iinv, w.p.collision_slop, bias_coef, inv_dt, g);
    } else geo_circles(A.px, A.py, A.r, B.px, B.py, B.r, A.b.minv, A.b.iinv, B.b.minv, B.b.iinv, w.p.collision_slop, bias_coef, inv_dt, g);
}

__device__ __forceinline__ bool body_fresh(const SapioWorld &w, uint32_t id, int NC, int ND, int32_t tick) {
    if ((int)id < NC) return w.org_birth_tick[id >> 6] == tick;
    if ((int)id < NC + ND) return w.d_state[id - NC] == 2;
    return false;
}

// per body: predecessor links of its contacts in canonical order (its adjacency order), contact index by construction
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
            if (c < nx) { if (y > (uint32_t)b) W.nx_prev_a[c] = prev; else W.nx_prev_b[c] = prev; }
            prev = c;
        }
    }
}

// per new cross contact: persistent-arbiter match (jnAcc/jtAcc of the previous step), bounce from pre-warm-start
// velocities (cpArbiterPreStep), coupling marks
__global__ void __launch_bounds__(128) k_cross_prestep(WORLD_ARG, Workspace W) {
    const SapioParams &p = w.p;
    const int NC = p.n_org * MAXC, ND = p.n_dead, WALL0 = NC + ND;
    const int nx = W.flags[FL_NX];
    const int32_t tick = (int32_t)w.g_tick[0];
    const float bc = bias_coef_of(p.collision_bias, p.dt_phys);
    const int pn = w.g_x_n[0];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nx; c += gridDim.x * blockDim.x) {
        uint64_t key = W.nx_key[c];
        uint32_t a = (uint32_t)(key >> 32), b = (uint32_t)key;
        GRef A, B; gload(w, a, NC, ND, A); gload(w, b, NC, ND, B);
        CGeo g; cross_geo(w, A, B, b, WALL0, bc, g);
        W.nx_bounce[c] = contact_bounce(A.b, B.b, g, A.e * B.e);
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

// ------------------------------------------------------------------------------------------------------------------
// organism solver: one warp per organism, state in shared memory
// ------------------------------------------------------------------------------------------------------------------
struct OrgS {
    float px[MAXC], py[MAXC], vx[MAXC], vy[MAXC], av[MAXC], vbx[MAXC], vby[MAXC], wb[MAXC];
    float minv[MAXC], iinv[MAXC], rad[MAXC], el[MAXC], fr[MAXC], relx[MAXC], rely[MAXC];
    float spx[MAXC], spy[MAXC], svx[MAXC], svy[MAXC], sjn[MAXC];
    float jn[SAPIO_NPAIR];
    float cb[SAPIO_ICAP], cjn[SAPIO_ICAP], cjt[SAPIO_ICAP], cjb[SAPIO_ICAP];
    uint16_t ccode[SAPIO_ICAP];
    uint8_t clev[SAPIO_ICAP];
    uint8_t last[MAXC];
    int nic, nlev;
};

struct OrgCtx {
    OrgS &s; const SapioWorld &w; int o, nc, lane;
    uint64_t alive, sens;
    float bc_contact, bc_joint, inv_dt;
    __device__ OrgCtx(OrgS &s_, const SapioWorld &w_, int o_) : s(s_), w(w_), o(o_), lane(lane_id()) {
        nc = w.org_ncells[o];
        bc_contact = bias_coef_of(w.p.collision_bias, w.p.dt_phys);
        bc_joint = bias_coef_of(w.p.joint_error_bias, w.p.dt_phys);
        inv_dt = 1.0f / w.p.dt_phys;
    }
    __device__ bool is_alive(int k) const { return (alive >> k) & 1ull; }

    // cells + sensor bodies + joint impulses -> shared memory
    __device__ void load() {
        uint32_t am[2] = {0, 0}, sm[2] = {0, 0};
        for (int h = 0; h < 2; h++) {
            int k = lane + 32 * h, row = o * MAXC + k;
            bool a = k < nc && w.c_alive[row] == SAPIO_CELL_ALIVE, sn = false;
            if (a) {
                float2 p = reinterpret_cast<const float2 *>(w.c_pos)[row], v = reinterpret_cast<const float2 *>(w.c_vel)[row];
                float2 vb = reinterpret_cast<const float2 *>(w.c_vbias)[row], rel = reinterpret_cast<const float2 *>(w.c_rel)[row];
                float m = w.c_mass[row], r = w.c_size[row] * 0.5f;
                s.px[k] = p.x; s.py[k] = p.y; s.vx[k] = v.x; s.vy[k] = v.y; s.av[k] = w.c_angvel[row];
                s.vbx[k] = vb.x; s.vby[k] = vb.y; s.wb[k] = w.c_wbias[row];
                s.minv[k] = 1.0f / m; s.iinv[k] = 1.0f / (0.5f * m * r * r); s.rad[k] = r;
                s.el[k] = w.c_elast[row]; s.fr[k] = w.c_fric[row]; s.relx[k] = rel.x; s.rely[k] = rel.y;
                sn = is_sensor_type(w.c_type[row]);
                if (sn) {
                    float2 sp = reinterpret_cast<const float2 *>(w.c_spos)[row], sv = reinterpret_cast<const float2 *>(w.c_svel)[row];
                    s.spx[k] = sp.x; s.spy[k] = sp.y; s.svx[k] = sv.x; s.svy[k] = sv.y; s.sjn[k] = w.c_spin_jn[row];
                }
            }
            am[h] = __ballot_sync(FULL, a); sm[h] = __ballot_sync(FULL, sn);
        }
        alive = ((uint64_t)am[1] << 32) | am[0]; sens = ((uint64_t)sm[1] << 32) | sm[0];
        const float *gj = w.j_jn + (size_t)o * SAPIO_NPAIR;
        for (int i = 0; i + 1 < nc; i++) {
            int base = pair_index(i, i + 1);
            for (int j = i + 1 + lane; j < nc; j += 32) s.jn[base + (j - i - 1)] = gj[base + (j - i - 1)];
        }
        __syncwarp();
    }

    // cpCollision: intra-organism cell/cell pairs and cell/wall pairs, in canonical (sorted code) order
    __device__ void detect() {
        int nic = 0;
        bool over = false;
        for (int i = 0; i < nc; i++) {
            if (!is_alive(i)) continue;
            float xi = s.px[i], yi = s.py[i], ri = s.rad[i];
            for (int base = 0; base < nc; base += 32) {
                int j = base + lane;
                bool hit = false;
                if (j > i && j < nc && is_alive(j)) { double d2; hit = circles_overlap(xi, yi, ri, s.px[j], s.py[j], s.rad[j], &d2); }
                uint32_t m = __ballot_sync(FULL, hit);
                if (hit) { int slot = nic + __popc(m & ((1u << lane) - 1)); if (slot < SAPIO_ICAP) s.ccode[slot] = (uint16_t)(i * 128 + j); else over = true; }
                nic += __popc(m);
            }
            // walls: cheap reject first (a cell farther than r + 2*frame from every border cannot touch)
            float fw2 = 2.0f * w.p.frame_width + ri + 1.0f;
            if (xi < fw2 || yi < fw2 || xi > w.p.width - fw2 || yi > w.p.height - fw2) {
                bool hit = false;
                if (lane < SAPIO_NWALL) { WallHit h; hit = circle_wall(w.p, lane, xi, yi, ri, &h); }
                uint32_t m = __ballot_sync(FULL, hit);
                if (hit) { int slot = nic + __popc(m & ((1u << lane) - 1)); if (slot < SAPIO_ICAP) s.ccode[slot] = (uint16_t)(i * 128 + MAXC + lane); else over = true; }
                nic += __popc(m);
            }
        }
        if (__any_sync(FULL, over) || nic > SAPIO_ICAP) { nic = SAPIO_ICAP; if (lane == 0) w.g_stats[SAPIO_STAT_OVERFLOW] = 1; }
        if (lane == 0) s.nic = nic;
        __syncwarp();
    }

    // persistent arbiters: jnAcc/jtAcc of the previous step's intra list (sorted by code)
    __device__ void match_prev() {
        const int pn = w.org_ic_n[o];
        const uint16_t *pc = w.ic_code + (size_t)o * SAPIO_ICAP;
        for (int c = lane; c < s.nic; c += 32) {
            uint16_t code = s.ccode[c];
            int lo = 0, hi = pn;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (pc[mid] < code) lo = mid + 1; else hi = mid; }
            bool f = lo < pn && pc[lo] == code;
            s.cjn[c] = f ? w.ic_jn[(size_t)o * SAPIO_ICAP + lo] : 0.f;
            s.cjt[c] = f ? w.ic_jt[(size_t)o * SAPIO_ICAP + lo] : 0.f;
            s.cjb[c] = 0.f;
        }
        __syncwarp();
    }

    // dependency levels: contact c runs after every earlier contact (list order) that shares a body with it
    __device__ void levels() {
        for (int k = lane; k < MAXC; k += 32) s.last[k] = 0;
        __syncwarp();
        if (lane == 0) {
            int mx = 0;
            for (int c = 0; c < s.nic; c++) {
                int i = s.ccode[c] >> 7, j = s.ccode[c] & 127;
                int L = s.last[i];
                if (j < MAXC && s.last[j] > L) L = s.last[j];
                L++;
                s.clev[c] = (uint8_t)L; s.last[i] = (uint8_t)L;
                if (j < MAXC) s.last[j] = (uint8_t)L;
                mx = max(mx, L);
            }
            s.nlev = mx;
        }
        __syncwarp();
    }

    __device__ void cbody(int k, CBody &b) const { b.vx = s.vx[k]; b.vy = s.vy[k]; b.w = s.av[k]; b.vbx = s.vbx[k]; b.vby = s.vby[k]; b.wb = s.wb[k]; b.minv = s.minv[k]; b.iinv = s.iinv[k]; }
    __device__ void cstore(int k, const CBody &b) { s.vx[k] = b.vx; s.vy[k] = b.vy; s.av[k] = b.w; s.vbx[k] = b.vbx; s.vby[k] = b.vby; s.wb[k] = b.wb; }
    __device__ void geo(int c, int &i, int &j, CGeo &g, float &e, float &mu) const {
        i = s.ccode[c] >> 7; j = s.ccode[c] & 127;
        if (j < MAXC) {
            geo_circles(s.px[i], s.py[i], s.rad[i], s.px[j], s.py[j], s.rad[j], s.minv[i], s.iinv[i], s.minv[j], s.iinv[j], w.p.collision_slop, bc_contact, inv_dt, g);
            e = s.el[i] * s.el[j]; mu = s.fr[i] * s.fr[j];
        } else {
            WallHit h; circle_wall(w.p, j - MAXC, s.px[i], s.py[i], s.rad[i], &h);
            geo_wall(s.px[i], s.py[i], s.rad[i], h, w.p.frame_width, s.minv[i], s.iinv[i], w.p.collision_slop, bc_contact, inv_dt, g);
            e = s.el[i] * w.p.wall_elast; mu = s.fr[i] * w.p.wall_fric;
        }
    }
    // cpArbiterPreStep: bounce from the velocities before any cached impulse is applied
    __device__ void prestep_contacts() {
        for (int c = lane; c < s.nic; c += 32) {
            int i, j; CGeo g; float e, mu;
            geo(c, i, j, g, e, mu);
            CBody A, B = CBody{0, 0, 0, 0, 0, 0, 0, 0};
            cbody(i, A); if (j < MAXC) cbody(j, B);
            s.cb[c] = contact_bounce(A, B, g, e);
        }
        __syncwarp();
    }
    // one sweep over the intra contacts in canonical order (levels); WARM: cached impulses instead of the solver step
    template <bool WARM>
    __device__ void contacts_pass(float dt_coef) {
        for (int L = 1; L <= s.nlev; L++) {
            for (int c = lane; c < s.nic; c += 32) {
                if (s.clev[c] != L) continue;
                int i, j; CGeo g; float e, mu;
                geo(c, i, j, g, e, mu);
                CBody A, B = CBody{0, 0, 0, 0, 0, 0, 0, 0};
                cbody(i, A); if (j < MAXC) cbody(j, B);
                if (WARM) contact_warm(A, B, g, s.cjn[c], s.cjt[c], dt_coef);
                else contact_apply(A, B, g, s.cb[c], mu, s.cjn[c], s.cjt[c], s.cjb[c]);
                cstore(i, A); if (j < MAXC) cstore(j, B);
            }
            __syncwarp();
        }
    }
    // cpPinJoint.c with anchors at the centres: sensor pins (sensor body -> cell, rest 0), then all pairs (i<j) in
    // lexicographic order evaluated as anti-diagonals t = i + j (pairs of one diagonal share no body)
    template <bool WARM>
    __device__ void joints_pass(float dt_coef) {
        for (int h = 0; h < 2; h++) {
            int k = lane + 32 * h;
            if (k < nc && ((sens >> k) & 1ull)) {
                float dx = s.px[k] - s.spx[k], dy = s.py[k] - s.spy[k];
                float d = sqrtf(dx * dx + dy * dy);
                float inv = d != 0.f ? 1.0f / d : 0.f, nx = dx * inv, ny = dy * inv;
                float mb = s.minv[k], nMass = 1.0f / (1.0f + mb);
                float jnv;
                if (WARM) jnv = s.sjn[k] * dt_coef;
                else {
                    float bias = -bc_joint * (d - 0.f) * inv_dt;
                    float vrn = (s.vx[k] - s.svx[k]) * nx + (s.vy[k] - s.svy[k]) * ny;
                    jnv = (bias - vrn) * nMass;
                    s.sjn[k] += jnv;
                }
                s.svx[k] -= nx * jnv; s.svy[k] -= ny * jnv;           // sensor body: mass 1
                s.vx[k] += nx * jnv * mb; s.vy[k] += ny * jnv * mb;
            }
        }
        __syncwarp();
        for (int t = 1; t <= 2 * nc - 3; t++) {
            int imin = max(0, t - (nc - 1)), imax = (t - 1) >> 1;
            int i = imin + lane, j = t - i;
            if (i <= imax && is_alive(i) && is_alive(j)) {
                float dx = s.px[j] - s.px[i], dy = s.py[j] - s.py[i];
                float d = sqrtf(dx * dx + dy * dy);
                float inv = d != 0.f ? 1.0f / d : 0.f, nx = dx * inv, ny = dy * inv;
                float ma = s.minv[i], mb = s.minv[j];
                int q = pair_index(i, j);
                float jnv;
                if (WARM) jnv = s.jn[q] * dt_coef;
                else {
                    float lx = s.relx[j] - s.relx[i], ly = s.rely[j] - s.rely[i];
                    float rest = sqrtf(lx * lx + ly * ly);
                    float bias = -bc_joint * (d - rest) * inv_dt;
                    float vrn = (s.vx[j] - s.vx[i]) * nx + (s.vy[j] - s.vy[i]) * ny;
                    jnv = (bias - vrn) * (1.0f / (ma + mb));
                    s.jn[q] += jnv;
                }
                float jx = nx * jnv, jy = ny * jnv;
                s.vx[i] -= jx * ma; s.vy[i] -= jy * ma;
                s.vx[j] += jx * mb; s.vy[j] += jy * mb;
            }
            __syncwarp();
        }
    }

    __device__ void store_vel() {
        for (int h = 0; h < 2; h++) {
            int k = lane + 32 * h, row = o * MAXC + k;
            if (k < nc && is_alive(k)) {
                reinterpret_cast<float2 *>(w.c_vel)[row] = make_float2(s.vx[k], s.vy[k]); w.c_angvel[row] = s.av[k];
                reinterpret_cast<float2 *>(w.c_vbias)[row] = make_float2(s.vbx[k], s.vby[k]); w.c_wbias[row] = s.wb[k];
                if ((sens >> k) & 1ull) { reinterpret_cast<float2 *>(w.c_svel)[row] = make_float2(s.svx[k], s.svy[k]); w.c_spin_jn[row] = s.sjn[k]; }
            }
        }
        float *gj = w.j_jn + (size_t)o * SAPIO_NPAIR;
        for (int i = 0; i + 1 < nc; i++) {
            int base = pair_index(i, i + 1);
            for (int j = i + 1 + lane; j < nc; j += 32) gj[base + (j - i - 1)] = s.jn[base + (j - i - 1)];
        }
    }
    __device__ void store_contacts(float *ws_bounce, float *ws_jbias) {
        for (int c = lane; c < s.nic; c += 32) {
            size_t g = (size_t)o * SAPIO_ICAP + c;
            w.ic_code[g] = s.ccode[c]; w.ic_jn[g] = s.cjn[c]; w.ic_jt[g] = s.cjt[c];
            if (ws_bounce) { ws_bounce[g] = s.cb[c]; ws_jbias[g] = s.cjb[c]; }
        }
        if (lane == 0) w.org_ic_n[o] = s.nic;
    }
    __device__ void load_contacts(const float *ws_bounce, const float *ws_jbias) {
        int n = w.org_ic_n[o];
        for (int c = lane; c < n; c += 32) {
            size_t g = (size_t)o * SAPIO_ICAP + c;
            s.ccode[c] = w.ic_code[g]; s.cjn[c] = w.ic_jn[g]; s.cjt[c] = w.ic_jt[g];
            s.cb[c] = ws_bounce[g]; s.cjb[c] = ws_jbias[g];
        }
        if (lane == 0) s.nic = n;
        __syncwarp();
    }
};

// organisms with no cross contact: the whole solve in one launch, everything on chip
#define ORG_WARPS 4
__global__ void __launch_bounds__(ORG_WARPS * 32) k_org_solve(WORLD_ARG, Workspace W) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OrgS &s = reinterpret_cast<OrgS *>(smem_raw)[threadIdx.x >> 5];
    const int warps = gridDim.x * ORG_WARPS;
    const float dt_coef = w.g_stepped[0] ? 1.0f : 0.0f;   // prev_dt == 0 on the very first step
    int nic_sum = 0;
    for (int o = blockIdx.x * ORG_WARPS + (threadIdx.x >> 5); o < w.p.n_org; o += warps) {
        if (w.org_state[o] != 1 || W.org_coupled[o]) continue;
        OrgCtx c(s, w, o);
        c.load();
        c.detect();
        c.match_prev();
        c.levels();
        c.prestep_contacts();
        if (dt_coef != 0.f) { c.contacts_pass<true>(dt_coef); c.joints_pass<true>(dt_coef); }
        for (int it = 0; it < w.p.iterations; it++) { c.contacts_pass<false>(0.f); c.joints_pass<false>(0.f); }
        c.store_vel();
        c.store_contacts(nullptr, nullptr);
        nic_sum += s.nic;
        __syncwarp();
    }
    if (lane_id() == 0 && nic_sum) atomicAdd(&W.flags[FL_NIC], nic_sum);
}

// ------------------------------------------------------------------------------------------------------------------
// coupled path: cooperative kernel
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_coupled_list(WORLD_ARG, Workspace W) {
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < w.p.n_org; o += gridDim.x * blockDim.x)
        if (w.org_state[o] == 1 && W.org_coupled[o]) W.coupled_list[atomicAdd(&W.flags[FL_NCOUPLED], 1)] = o;
}

template <bool WARM>
__device__ __forceinline__ void cross_level_pass(const SapioWorld &w, const Workspace &W, int L, int nx, float dt_coef, int gtid, int gthreads) {
    const int NC = w.p.n_org * MAXC, ND = w.p.n_dead, WALL0 = NC + ND;
    const float bc = bias_coef_of(w.p.collision_bias, w.p.dt_phys);
    for (int c = gtid; c < nx; c += gthreads) {
        if (W.nx_level[c] != L) continue;
        uint64_t key = W.nx_key[c];
        uint32_t a = (uint32_t)(key >> 32), b = (uint32_t)key;
        GRef A, B; gload(w, a, NC, ND, A); gload(w, b, NC, ND, B);
        CGeo g; cross_geo(w, A, B, b, WALL0, bc, g);
        if (WARM) contact_warm(A.b, B.b, g, W.nx_jn[c], W.nx_jt[c], dt_coef);
        else { float jn = W.nx_jn[c], jt = W.nx_jt[c], jb = W.nx_jbias[c]; contact_apply(A.b, B.b, g, W.nx_bounce[c], A.u * B.u, jn, jt, jb); W.nx_jn[c] = jn; W.nx_jt[c] = jt; W.nx_jbias[c] = jb; }
        gstore(w, a, NC, A); gstore(w, b, NC, B);
    }
}

__global__ void __launch_bounds__(ORG_WARPS * 32) k_coupled(WORLD_ARG, Workspace W) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OrgS &s = reinterpret_cast<OrgS *>(smem_raw)[threadIdx.x >> 5];
    const int nx = W.flags[FL_NX];
    if (nx == 0) return;                                       // uniform over the grid: no barrier has been entered
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    const int gwarp = gtid >> 5, gwarps = gthreads >> 5;
    const int ncoupled = W.flags[FL_NCOUPLED];
    const float dt_coef = w.g_stepped[0] ? 1.0f : 0.0f;
    // dependency levels: level[c] = 1 + max(level of the previous contact on either body); fixed point == longest chain
    for (;;) {
        if (gtid == 0) W.flags[FL_CHANGED] = 0;
        grid.sync();
        bool ch = false;
        for (int c = gtid; c < nx; c += gthreads) {
            int pa = W.nx_prev_a[c], pb = W.nx_prev_b[c];
            int L = 1 + max(pa >= 0 ? W.nx_level[pa] : 0, pb >= 0 ? W.nx_level[pb] : 0);
            if (L > W.nx_level[c]) { W.nx_level[c] = L; ch = true; atomicMax(&W.flags[FL_MAXLEVEL], L); }
        }
        if (ch) W.flags[FL_CHANGED] = 1;
        grid.sync();
        int any = W.flags[FL_CHANGED];
        grid.sync();
        if (!any) break;
    }
    const int maxL = max(W.flags[FL_MAXLEVEL], 1);
    // stage A: coupled organisms detect their intra contacts and take their bounce before any velocity changes
    for (int q = gwarp; q < ncoupled; q += gwarps) {
        OrgCtx c(s, w, W.coupled_list[q]);
        c.load(); c.detect(); c.match_prev(); c.prestep_contacts();
        c.store_contacts(W.ic_bounce, W.ic_jbias);
        if (lane_id() == 0) atomicAdd(&W.flags[FL_NIC], s.nic);
        __syncwarp();
    }
    grid.sync();
    // stage B/C: cached impulses -- cross contacts by level, then each coupled organism's own arbiters and constraints
    if (dt_coef != 0.f) {
        for (int L = 1; L <= maxL; L++) { cross_level_pass<true>(w, W, L, nx, dt_coef, gtid, gthreads); grid.sync(); }
        for (int q = gwarp; q < ncoupled; q += gwarps) {
            OrgCtx c(s, w, W.coupled_list[q]);
            c.load(); c.load_contacts(W.ic_bounce, W.ic_jbias); c.levels();
            c.contacts_pass<true>(dt_coef); c.joints_pass<true>(dt_coef);
            c.store_vel();
            __syncwarp();
        }
        grid.sync();
    }
    for (int it = 0; it < w.p.iterations; it++) {
        for (int L = 1; L <= maxL; L++) { cross_level_pass<false>(w, W, L, nx, 0.f, gtid, gthreads); grid.sync(); }
        for (int q = gwarp; q < ncoupled; q += gwarps) {
            OrgCtx c(s, w, W.coupled_list[q]);
            c.load(); c.load_contacts(W.ic_bounce, W.ic_jbias); c.levels();
            c.contacts_pass<false>(0.f); c.joints_pass<false>(0.f);
            c.store_vel(); c.store_contacts(W.ic_bounce, W.ic_jbias);
            __syncwarp();
        }
        grid.sync();
    }
}

// publish: the new arbiter list becomes w.x_*, the adjacency w.xa_*; created-this-tick marks are cleared
__global__ void __launch_bounds__(256) k_step_finish(WORLD_ARG, Workspace W) {
    const int NC = w.p.n_org * MAXC, ND = w.p.n_dead, NB = NC + ND;
    const int nx = W.flags[FL_NX], na = W.flags[FL_NADJ];
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    for (int c = gtid; c < nx; c += gthreads) { w.x_key[c] = W.nx_key[c]; w.x_jn[c] = W.nx_jn[c]; w.x_jt[c] = W.nx_jt[c]; }
    for (int q = gtid; q < na; q += gthreads) w.xa_other[q] = W.adj_other[q];
    for (int b = gtid; b <= NB; b += gthreads) w.xa_off[b] = min(W.off_all[b], na);
    for (int d = gtid; d < ND; d += gthreads) if (w.d_state[d] == 2) w.d_state[d] = 1;
    if (gtid == 0) {
        w.g_x_n[0] = nx; w.g_stepped[0] = 1;
        w.g_stats[SAPIO_STAT_XCONTACTS] = nx; w.g_stats[SAPIO_STAT_ICONTACTS] = W.flags[FL_NIC];
        w.g_stats[SAPIO_STAT_LEVELS] = W.flags[FL_MAXLEVEL]; w.g_stats[SAPIO_STAT_COUPLED] = W.flags[FL_NCOUPLED];
        if (W.flags[FL_OVERFLOW]) w.g_stats[SAPIO_STAT_OVERFLOW] += 1;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host-side launch sequence
// ------------------------------------------------------------------------------------------------------------------
static int g_coop_blocks = 0;

static int launch_space_step(const SapioWorld &w, const Workspace &W, cudaStream_t st, char *err, size_t errlen) {
#define STEP_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { snprintf(err, errlen, #expr ": %s", cudaGetErrorString(e_)); return SAPIO_E_CUDA; } } while (0)
    const SapioParams &p = w.p;
    const int NC = p.n_org * MAXC, NB = NC + p.n_dead, NG = p.grid_nx * p.grid_ny;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto sgrid = [&](long items, int threads) { long want = (items + threads - 1) / threads, wave = (long)sms * 8; return (int)(want < 1 ? 1 : (want > wave ? wave : want)); };

    STEP_TRY(cudaMemsetAsync(W.flags, 0, 256, st));
    STEP_TRY(cudaMemsetAsync(W.org_coupled, 0, (size_t)p.n_org, st));
    k_integrate<<<sgrid(NB, 256), 256, 0, st>>>(w, W.gkey, W.gval);
    int bits = 1; while ((1ll << bits) < (long long)NG + 1 && bits < 32) bits++;
    size_t tmp = W.cub_bytes;
    // inactive slots carry key ~0: with end_bit = bits they sort by their low bits only, so mask them out by sorting all 32
    STEP_TRY(cub::DeviceRadixSort::SortPairs(W.cub_tmp, tmp, W.gkey, W.skey, W.gval, W.sval, NB, 0, 32, st));
    (void)bits;
    STEP_TRY(cudaMemsetAsync(W.cell_start, 0, (size_t)NG * 4, st));
    STEP_TRY(cudaMemsetAsync(W.cell_end, 0, (size_t)NG * 4, st));
    k_cell_bounds<<<sgrid(NB, 256), 256, 0, st>>>(W.skey, NB, NG, W.cell_start, W.cell_end);
    k_narrow<false><<<sgrid(NB + 1, 128), 128, 0, st>>>(w, W);
    tmp = W.cub_bytes;
    STEP_TRY(cub::DeviceScan::ExclusiveSum(W.cub_tmp, tmp, W.cnt_up, W.off_up, NB + 1, st));
    tmp = W.cub_bytes;
    STEP_TRY(cub::DeviceScan::ExclusiveSum(W.cub_tmp, tmp, W.cnt_all, W.off_all, NB + 1, st));
    k_narrow_totals<<<1, 1, 0, st>>>(w, W);
    k_narrow<true><<<sgrid(NB + 1, 128), 128, 0, st>>>(w, W);
    STEP_TRY(cudaMemsetAsync(W.nx_prev_a, 0xFF, (size_t)p.x_cap * 4, st));
    STEP_TRY(cudaMemsetAsync(W.nx_prev_b, 0xFF, (size_t)p.x_cap * 4, st));
    k_cross_links<<<sgrid(NB, 128), 128, 0, st>>>(w, W);
    k_cross_prestep<<<sgrid(p.x_cap, 128), 128, 0, st>>>(w, W);
    k_coupled_list<<<sgrid(p.n_org, 256), 256, 0, st>>>(w, W);

    const size_t smem = sizeof(OrgS) * ORG_WARPS;
    static bool attr_done = false;
    if (!attr_done) {
        STEP_TRY(cudaFuncSetAttribute(k_org_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        STEP_TRY(cudaFuncSetAttribute(k_coupled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        STEP_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_coupled, ORG_WARPS * 32, smem));
        g_coop_blocks = per_sm * sms;
        attr_done = true;
    }
    {   // one warp per organism; a whole number of CTAs per SM
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_org_solve, ORG_WARPS * 32, smem);
        long want = ((long)p.n_org + ORG_WARPS - 1) / ORG_WARPS, wave = (long)sms * (per_sm > 0 ? per_sm : 1);
        long waves = (want + wave - 1) / wave;
        int blocks = (int)(want <= wave ? want : wave * (waves > 8 ? 8 : waves));
        if (blocks < 1) blocks = 1;
        k_org_solve<<<blocks, ORG_WARPS * 32, smem, st>>>(w, W);
    }
    {
        if (g_coop_blocks < 1) { snprintf(err, errlen, "k_coupled does not fit on an SM"); return SAPIO_E_CUDA; }
        SapioWorld wc = w; Workspace Wc = W;
        void *args[] = { (void *)&wc, (void *)&Wc };
        STEP_TRY(cudaLaunchCooperativeKernel((void *)k_coupled, dim3(g_coop_blocks), dim3(ORG_WARPS * 32), args, smem, st));
    }
    k_step_finish<<<sgrid(NB + 1, 256), 256, 0, st>>>(w, W);
    STEP_TRY(cudaGetLastError());
    return 0;
#undef STEP_TRY
}
