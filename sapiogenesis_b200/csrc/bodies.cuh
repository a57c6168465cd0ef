// bodies.cuh -- per-body streaming kernels of Space.step / updateSprite: position integration, friction, bookkeeping.
// HBM-bound: one thread per body row, float2-vectorised, every array touched once.
#pragma once
#include "common.cuh"

// cpBodyUpdatePosition (Chipmunk cpBody.c; SURVEY A-2 step 1) for every dynamic body: living cells, their sensor
// bodies (sprites.py:300-323) and free dead cells. p = fma(v + v_bias, dt, p): the single fp32 op sequence the oracle mirrors.
// Also writes the broadphase key of every solid body (grid cell index, or ~0 for "no body in this slot").
#define INTEGRATE_STAGE 256
template <bool INTEGRATE>
__global__ void __launch_bounds__(256) k_integrate(WORLD_ARG, uint32_t *__restrict__ grid_key, uint32_t *__restrict__ live_list, int *__restrict__ cell_cnt, int *__restrict__ n_live) {
    const SapioParams &p = w.p;
    const int NC = p.n_org * MAXC, NB = NC + p.n_dead;
    const float dt = p.dt_phys;
    // the list of living bodies is staged per warp in shared memory and appended in runs of ~200 entries: one atomic on the list's counter
    // per run instead of one per warp and iteration (300k atomics on ONE address per tick were most of this kernel's time)
    __shared__ uint32_t stage[8][INTEGRATE_STAGE];
    uint32_t *mine = stage[threadIdx.x >> 5];
    int staged = 0;                                          // warp-uniform among the lanes still in the loop (lane 0 leaves last)
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < NB; b += gridDim.x * blockDim.x) {
        uint32_t key = 0xFFFFFFFFu;
        float2 pos;
        bool solid = false;
        if (b < NC) {
            int o = b >> 6;
            if (w.org_state[o] == 1 && (b & 63) < w.org_ncells[o] && w.c_alive[b] == SAPIO_CELL_ALIVE) {
                float2 *P = reinterpret_cast<float2 *>(w.c_pos), *V = reinterpret_cast<float2 *>(w.c_vel), *VB = reinterpret_cast<float2 *>(w.c_vbias);
                pos = P[b];
                if (INTEGRATE) {
                float2 v = V[b], vb = VB[b];
                pos.x = __fmaf_rn(__fadd_rn(v.x, vb.x), dt, pos.x);
                pos.y = __fmaf_rn(__fadd_rn(v.y, vb.y), dt, pos.y);
                P[b] = pos; VB[b] = make_float2(0.f, 0.f);
                w.c_ang[b] = __fmaf_rn(__fadd_rn(w.c_angvel[b], w.c_wbias[b]), dt, w.c_ang[b]);
                w.c_wbias[b] = 0.f;
                }
                if (INTEGRATE && is_sensor_type(w.c_type[b])) {
                    float2 *SP = reinterpret_cast<float2 *>(w.c_spos), *SV = reinterpret_cast<float2 *>(w.c_svel), *SVB = reinterpret_cast<float2 *>(w.c_svbias);
                    float2 sp = SP[b], sv = SV[b], svb = SVB[b];
                    sp.x = __fmaf_rn(__fadd_rn(sv.x, svb.x), dt, sp.x);
                    sp.y = __fmaf_rn(__fadd_rn(sv.y, svb.y), dt, sp.y);
                    SP[b] = sp; SVB[b] = make_float2(0.f, 0.f);
                }
                solid = true;
            }
        } else {
            int d = b - NC;
            if (w.d_state[d]) {
                float2 *P = reinterpret_cast<float2 *>(w.d_pos), *V = reinterpret_cast<float2 *>(w.d_vel), *VB = reinterpret_cast<float2 *>(w.d_vbias);
                pos = P[d];
                if (INTEGRATE) {
                float2 v = V[d], vb = VB[d];
                pos.x = __fmaf_rn(__fadd_rn(v.x, vb.x), dt, pos.x);
                pos.y = __fmaf_rn(__fadd_rn(v.y, vb.y), dt, pos.y);
                P[d] = pos; VB[d] = make_float2(0.f, 0.f);
                w.d_ang[d] = __fmaf_rn(__fadd_rn(w.d_angvel[d], w.d_wbias[d]), dt, w.d_ang[d]);
                w.d_wbias[d] = 0.f;
                }
                solid = true;
            }
        }
        // Only living bodies have a grid key, a place in their cell's count and in the list (its consumers -- k_grid_scatter / k_grid_rank --
        // read the key through the list): two thirds of the slots are empty, most of them in runs longer than a warp (the unused rows of an
        // organism, the free tail of the dead pool), and those warps leave here after one ballot.
        if (grid_key) {
            const int left = NB - (b - (int)(threadIdx.x & 31));             // the lanes of this warp that are in this iteration: slots b0 .. b0 + 31 below NB
            const uint32_t act = left >= 32 ? FULL : ((1u << left) - 1u);
            const uint32_t m = __ballot_sync(act, solid);
            if (solid) {
                // same cell function as the oracle's candidate grid; any superset generator is exact
                int cx = (int)floorf((pos.x - p.grid_x0) / p.grid_cell) + 1, cy = (int)floorf(pos.y / p.grid_cell) + 1;
                cx = min(max(cx, 0), p.grid_nx - 1); cy = min(max(cy, 0), p.grid_ny - 1);
                key = (uint32_t)(cy * p.grid_nx + cx);
                grid_key[b] = key;
                // the cells of an organism share a handful of grid cells and sit in consecutive slots: one atomic per (warp, cell), not per body
                const uint32_t peers = __match_any_sync(m, key);
                const int lane = threadIdx.x & 31;
                if (lane == __ffs(peers) - 1) atomicAdd(&cell_cnt[key], __popc(peers));
                // the living bodies as a list (order immaterial: grid.cuh ranks them inside their cells)
                mine[staged + __popc(m & ((1u << lane) - 1u))] = (uint32_t)b;
            }
            staged += __popc(m);
            if (staged > INTEGRATE_STAGE - 32) {             // flush: every lane of this iteration takes part (lanes 0 .. popc(act) - 1)
                __syncwarp(act);
                const int lane = threadIdx.x & 31;
                int base = 0;
                if (lane == 0) base = atomicAdd(n_live, staged);
                base = __shfl_sync(act, base, 0);
                for (int i = lane; i < staged; i += __popc(act)) live_list[base + i] = mine[i];
                __syncwarp(act);
                staged = 0;
            }
        }
    }
    if (grid_key) {                                          // the rest (lane 0 was in every iteration of its warp: its count is the warp's)
        __syncwarp();
        staged = __shfl_sync(FULL, staged, 0);
        if (staged) {
            const int lane = threadIdx.x & 31;
            int base = 0;
            if (lane == 0) base = atomicAdd(n_live, staged);
            base = __shfl_sync(FULL, base, 0);
            for (int i = lane; i < staged; i += 32) live_list[base + i] = mine[i];
        }
    }
}

// Sprite.updateSprite (physics.py:250-275, SURVEY E-8): v -= friction% * v * uDiff, same for omega; solid sprites only
// (living cells and dead cells), never the sensor bodies.
__global__ void __launch_bounds__(256) k_friction(WORLD_ARG) {
    const SapioParams &p = w.p;
    const int NC = p.n_org * MAXC, NB = NC + p.n_dead;
    const float f = p.friction_pct / 100.0f, dt = p.dt_wall;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < NB; b += gridDim.x * blockDim.x) {
        if (b < NC) {
            int o = b >> 6;
            if (w.org_state[o] != 1 || (b & 63) >= w.org_ncells[o] || w.c_alive[b] != SAPIO_CELL_ALIVE) continue;
            float2 *V = reinterpret_cast<float2 *>(w.c_vel);
            float2 v = V[b];
            v.x = v.x - f * v.x * dt; v.y = v.y - f * v.y * dt;
            V[b] = v;
            float a = w.c_angvel[b];
            w.c_angvel[b] = a - f * a * dt;
        } else {
            int d = b - NC;
            if (!w.d_state[d]) continue;
            float2 *V = reinterpret_cast<float2 *>(w.d_vel);
            float2 v = V[d];
            v.x = v.x - f * v.x * dt; v.y = v.y - f * v.y * dt;
            V[d] = v;
            float a = w.d_angvel[d];
            w.d_angvel[d] = a - f * a * dt;
        }
    }
}

// updateUI (Sapiogenesis.py:57-65): clamp gases at zero, recount the population; plus the tick counter.
__global__ void k_post_count(WORLD_ARG, int *__restrict__ scratch_pop) {
    int n = 0;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < w.p.n_org; o += gridDim.x * blockDim.x) n += (w.org_state[o] == 1 && !(w.org_flags[o] & SAPIO_F_GHOST));
    for (int s = 16; s; s >>= 1) n += __shfl_xor_sync(FULL, n, s);
    if (lane_id() == 0 && n) atomicAdd(scratch_pop, n);
}
__global__ void k_post_finish(WORLD_ARG, int *__restrict__ scratch_pop, int do_post, int advance) {
    if (do_post) {
        if (w.g_co2[0] < 0) w.g_co2[0] = 0;
        if (w.g_o2[0] < 0) w.g_o2[0] = 0;
        w.g_population[0] = *scratch_pop;
    }
    if (advance) w.g_tick[0] += 1;
}
