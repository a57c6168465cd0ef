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

#include "steptypes.cuh"
#include "grid.cuh"
#include "solver_api.h"
// ------------------------------------------------------------------------------------------------------------------
// broadphase
// ------------------------------------------------------------------------------------------------------------------
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
    const int n_live = W.flags[FL_NLIVE];
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_live; idx += gridDim.x * blockDim.x) {
        const uint32_t key = W.skey[idx];
        const float4 me = W.srec[idx];
        const int b = __float_as_int(me.w);
        const float mx = me.x, my = me.y, mr = me.z;
        const int my_org = b < NC ? (b >> 6) : -1;
        const int cx = (int)(key % (uint32_t)p.grid_nx), cy = (int)(key / (uint32_t)p.grid_nx);
        uint32_t loc[NARROW_LOCAL];
        int n = 0, up = 0;
        bool over = false;
        for (int yy = cy - 1; yy <= cy + 1; yy++) {
            if (yy < 0 || yy >= p.grid_ny) continue;
            // the three cells of a grid row are neighbours in the cell-ordered list: one run of records
            const int c0 = yy * p.grid_nx + max(cx - 1, 0), c1 = yy * p.grid_nx + min(cx + 1, p.grid_nx - 1);
            for (int i = W.cell_start[c0], e = W.cell_start[c1 + 1]; i < e; i++) {
                const float4 ot = W.srec[i];
                const int j = __float_as_int(ot.w);
                if (j == b) continue;
                if (my_org >= 0 && j < NC && (j >> 6) == my_org) continue;
                double d2;
                if (!circles_overlap(mx, my, mr, ot.x, ot.y, ot.z, &d2)) continue;
                if (n < NARROW_LOCAL) { loc[n] = (uint32_t)j; n++; up += (j > b); } else over = true;
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

// organisms -> two lists sorted by LIVING cells (counting sort: histogram, scan, scatter): coupled organisms and the rest. The solver
// works in rank space (living cells compacted), so what sizes an organism's slab, its lanes and its chain of wavefronts is the number
// of living cells, not the rows it occupies (an old organism carries the markers of the cells it lost).
__global__ void k_classify_count(WORLD_ARG, Workspace W, int dbg) {
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < w.p.n_org; o += gridDim.x * blockDim.x) {
        if (w.org_state[o] != 1) continue;
        const int rows = min(max(w.org_ncells[o], 0), MAXC);
        const uint4 *al = reinterpret_cast<const uint4 *>(w.c_alive + (size_t)o * MAXC);          // 64 bytes per organism, 16-byte aligned
        int live = 0;
#pragma unroll
        for (int q = 0; q < MAXC / 16; q++) {
            if (q * 16 >= rows) break;
            const uint4 v = al[q];
            const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int bb = 0; bb < 4; bb++) { const int k = q * 16 + u * 4 + bb; live += (k < rows && ((wds[u] >> (8 * bb)) & 255u) == SAPIO_CELL_ALIVE); }
        }
#ifdef SAPIO_DEBUG_FLAGS                               /* A/B switch 1: round-1 classes (rows in use) */
        if (dbg & 1) live = rows;
#endif
        W.org_nlive[o] = (uint8_t)live;
        const int nc = min(max(live, 1), MAXC);
        if (W.org_coupled[o]) { atomicAdd(&W.flags[FL_NCOUPLED], 1); atomicAdd(&W.flags[FL_CHIST + nc], 1); }
        else atomicAdd(&W.size_hist[nc], 1);
    }
}
__global__ void k_classify_scan(Workspace W) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int acc = 0, cacc = 0;
    for (int nc = 1; nc <= MAXC; nc++) {
        if (nc == 1 || ((nc - 1) & 7) == 0) { W.flags[FL_CLS0 + org_class_of(nc)] = acc; W.flags[FL_CCLS0 + org_class_of(nc)] = cacc; }   // first size of a class
        int c = W.size_hist[nc]; W.size_hist[nc] = acc; acc += c;
        c = W.flags[FL_CHIST + nc]; W.flags[FL_CHIST + nc] = cacc; cacc += c;
    }
    W.flags[FL_CLS0 + ORG_CLASSES] = acc; W.flags[FL_CCLS0 + ORG_CLASSES] = cacc;
}
__global__ void k_classify_scatter(WORLD_ARG, Workspace W) {
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < w.p.n_org; o += gridDim.x * blockDim.x) {
        if (w.org_state[o] != 1) continue;
        const int nc = min(max((int)W.org_nlive[o], 1), MAXC);
        if (W.org_coupled[o]) W.coupled_list[atomicAdd(&W.flags[FL_CHIST + nc], 1)] = o;
        else W.sorted_orgs[atomicAdd(&W.size_hist[nc], 1)] = o;
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
        w.g_stats[SAPIO_STAT_RARE_PATHS] |= W.flags[FL_RARE];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host-side launch sequence
// ------------------------------------------------------------------------------------------------------------------
struct StepPlan { bool ready; int sms; int simple_blocks; size_t simple_smem; int comp_blocks[COMP_TIERS_N]; int org_blocks[ORG_CLASSES]; size_t org_smem[ORG_CLASSES]; cudaStream_t side[ORG_CLASSES + COMP_TIERS_N]; cudaEvent_t fork, join[ORG_CLASSES + COMP_TIERS_N]; };
static StepPlan g_plan = {};
static int g_debug_flags = 0;                             // sapio_debug_flags (diagnostic builds only)
static int g_force_fallback = 0;                          // sapio_set_coupled_pack_from(INT_MAX): tests compare the two implementations

static int launch_space_step(const SapioWorld &w, const Workspace &W, cudaStream_t st, char *err, size_t errlen) {
#define STEP_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { snprintf(err, errlen, #expr ": %s", cudaGetErrorString(e_)); return SAPIO_E_CUDA; } } while (0)
    const SapioParams &p = w.p;
    const int NC = p.n_org * MAXC, NB = NC + p.n_dead, NG = p.grid_nx * p.grid_ny;
    if (!g_plan.ready) {
        int dev = 0;
        STEP_TRY(cudaGetDevice(&dev));
        STEP_TRY(cudaDeviceGetAttribute(&g_plan.sms, cudaDevAttrMultiProcessorCount, dev));
        for (int c = 0; c < ORG_CLASSES; c++) STEP_TRY(ORG_PLAN[c](g_plan.sms, &g_plan.org_blocks[c], &g_plan.org_smem[c]));
        STEP_TRY(cudaEventCreateWithFlags(&g_plan.fork, cudaEventDisableTiming));
        for (int c = 0; c < ORG_CLASSES + COMP_TIERS_N; c++) {
            STEP_TRY(cudaStreamCreateWithFlags(&g_plan.side[c], cudaStreamNonBlocking));
            STEP_TRY(cudaEventCreateWithFlags(&g_plan.join[c], cudaEventDisableTiming));
        }
        STEP_TRY(coupled_plan(g_plan.sms, &g_plan.simple_blocks, &g_plan.simple_smem));
        if (g_plan.simple_blocks < 1) { snprintf(err, errlen, "the coupled fallback kernel does not fit on an SM"); return SAPIO_E_CUDA; }
        STEP_TRY(component_plan(g_plan.sms, g_plan.comp_blocks));
        g_plan.ready = true;
    }
    const int sms = g_plan.sms;
    auto sgrid = [&](long items, int threads) { long want = (items + threads - 1) / threads, wave = (long)sms * 8; return (int)(want < 1 ? 1 : (want > wave ? wave : want)); };
    StepK K;
    K.dt = (double)p.dt_phys;
    K.bc_contact = 1.0 - pow((double)p.collision_bias, K.dt);        // cpSpaceStep: biasCoef = 1 - collisionBias^dt
    K.bc_joint = 1.0 - pow((double)p.joint_error_bias, K.dt);        // bias_coef(errorBias, dt)
    K.dbg = g_debug_flags;

    prof_mark(st, PT_BROADPHASE);
    STEP_TRY(cudaMemsetAsync(W.flags, 0, FL_BYTES, st));
    STEP_TRY(cudaMemsetAsync(W.org_coupled, 0, (size_t)p.n_org, st));
    STEP_TRY(cudaMemsetAsync(W.cell_cnt, 0, (size_t)(NG + 1) * 4, st));
    k_integrate<true><<<sgrid(NB, 256), 256, 0, st>>>(w, W.gkey, W.gval, W.cell_cnt, W.flags + FL_NLIVE);
    scan_exclusive_i32(W.cell_cnt, W.cell_start, NG + 1, W.scan_parts, st);
    k_grid_scatter<<<sgrid(NB / 3, 256), 256, 0, st>>>(W, W.flags + FL_NLIVE);
    k_grid_rank<<<sgrid(NB / 3, 256), 256, 0, st>>>(w, W, W.flags + FL_NLIVE);
    prof_mark(st, PT_NARROW);
    STEP_TRY(cudaMemsetAsync(W.cnt_up, 0, (size_t)(NB + 1) * 4, st));
    STEP_TRY(cudaMemsetAsync(W.cnt_all, 0, (size_t)(NB + 1) * 4, st));
    k_narrow_find<<<sgrid(NB / 3, 128), 128, 0, st>>>(w, W);
    scan_exclusive_i32(W.cnt_up, W.off_up, NB + 1, W.scan_parts, st);
    scan_exclusive_i32(W.cnt_all, W.off_all, NB + 1, W.scan_parts, st);
    k_narrow_totals<<<1, 1, 0, st>>>(w, W);
    k_narrow_fill<<<sgrid(2 * (long)p.x_cap, 128), 128, 0, st>>>(w, W);
    STEP_TRY(cudaMemsetAsync(W.nx_prev_a, 0xFF, (size_t)p.x_cap * 4, st));
    STEP_TRY(cudaMemsetAsync(W.nx_prev_b, 0xFF, (size_t)p.x_cap * 4, st));
    k_cross_links<<<sgrid(NB, 128), 128, 0, st>>>(w, W);
    k_cross_prestep<<<sgrid(p.x_cap, 128), 128, 0, st>>>(w, W, K);
    k_classify_count<<<sgrid(p.n_org, 256), 256, 0, st>>>(w, W, K.dbg);
    k_classify_scan<<<1, 32, 0, st>>>(W);
    k_classify_scatter<<<sgrid(p.n_org, 256), 256, 0, st>>>(w, W);
    LAUNCHES(20);
    // connected components of the cross-contact graph (components.cuh): planned here, solved next to the organisms without contacts
    prof_mark(st, PT_COMP_PLAN);
    component_plan_launch(w, W, g_force_fallback, sgrid(p.x_cap, 256), st);
    LAUNCHES(5);
    prof_mark(st, PT_ORG_SOLVE);
    {   // Persistent CTAs per organism size class and per component tier, all independent of each other, on side streams: every one
        // of them is a chain of dependent steps per organism whose throughput is set by how many are resident, so the tail of one
        // launch is filled by the next. Longest chains first.
        STEP_TRY(cudaEventRecord(g_plan.fork, st));
        int k = 0;
        auto side_begin = [&]() -> cudaStream_t { cudaStream_t ss = g_plan.side[k]; cudaStreamWaitEvent(ss, g_plan.fork, 0); return ss; };
        auto side_end = [&]() { cudaEventRecord(g_plan.join[k], g_plan.side[k]); cudaStreamWaitEvent(st, g_plan.join[k], 0); k++; };
        auto org = [&](int cls) { cudaStream_t ss = side_begin(); ORG_LAUNCH_FN[cls](w, W, K, cls, g_plan.org_blocks[cls], g_plan.org_smem[cls], ss); side_end(); };
        auto comp = [&](int t) { cudaStream_t ss = side_begin(); component_launch(w, W, K, t, g_plan.comp_blocks[t], ss); side_end(); };
        comp(6); comp(5); comp(4); org(7); org(6); comp(3); org(5); org(4); comp(2); org(3); org(2); comp(1); org(1); org(0); comp(0);
    }
    LAUNCHES(8 + COMP_TIERS_N);
    // the cooperative kernel only has work (decided on the device) when a component exceeds the largest tier
    prof_mark(st, PT_COUPLED);
    STEP_TRY(coupled_launch(w, W, K, g_plan.simple_blocks < 64 ? g_plan.simple_blocks : 64, g_plan.simple_smem, st));   // a small grid: its barriers are cheap, its work is rare
    LAUNCHES(1);
    prof_mark(st, PT_STEP_FINISH);
    k_step_finish<<<sgrid(NB + 1, 256), 256, 0, st>>>(w, W, nullptr);
    LAUNCHES(2);
    prof_mark(st, PT_END);
    STEP_TRY(cudaGetLastError());
    return 0;
#undef STEP_TRY
}

