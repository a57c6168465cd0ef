// frame.cuh -- what a GUI host exchanges with the world every tick (the end-to-end path of sapio_tick_host):
//   in : the event block updateUI applies (drought / poison amounts, Sapiogenesis.py:37-47)
//   out: the counters updateUI shows (Sapiogenesis.py:62-69) and the pose of every solid sprite, packed
//        (Sprite.setPos / setRotation, physics.py:253-272)
#pragma once
#include "common.cuh"
#include "rng.cuh"

struct FrameWs { SapioControl *control; SapioFrameHeader *header; float4 *cells; float4 *dead; int *counts; };

// updateUI world events, applied after the physics step like the reference's postUpdateEvent
__global__ void __launch_bounds__(256) k_events(WORLD_ARG, const SapioControl *__restrict__ ctl) {
    const float poison = ctl->poison, drought = ctl->drought;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    if (poison != 0.f) {
        const int NC = w.p.n_org * MAXC;
        for (int row = gtid; row < NC; row += gridDim.x * blockDim.x) {
            int o = row >> 6;
            if (w.org_state[o] == 1 && (row & 63) < w.org_ncells[o]) w.c_health[row] -= poison * 5.0f;     // Sapiogenesis.py:42-47
        }
    }
    if (gtid == 0 && drought != 0.f) { w.g_o2[0] -= (double)drought; w.g_co2[0] -= (double)drought; }   // Sapiogenesis.py:38-41
}

// random dead cell event (Sapiogenesis.py:48-53): with probability `random_cells` one piece of food (add_dead_clicked,
// Sapiogenesis.py:490-507) appears at (randrange(10, int(width)), randrange(10, int(height))), in the lowest free dead slot
__global__ void __launch_bounds__(256) k_event_food(WORLD_ARG, const SapioControl *__restrict__ ctl) {
    const float rate = ctl->random_cells;
    if (!(rate > 0.f)) return;
    uint32_t r[4];
    rng4(w.p.seed, -1, (int32_t)w.g_tick[0], RNG_EVENTS, 0, r);
    if (!(u01(r[0]) <= rate)) return;
    __shared__ int s_free;
    if (threadIdx.x == 0) s_free = 1 << 30;
    __syncthreads();
    for (int d = threadIdx.x; d < w.p.n_dead; d += blockDim.x) {
        if (d >= s_free) break;
        if (w.d_state[d] == 0) { atomicMin(&s_free, d); break; }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int d = s_free;
    if (d >= w.p.n_dead) { stat_add(w, SAPIO_STAT_OVERFLOW, 1); return; }
    const int wi = (int)w.p.width, hi = (int)w.p.height;
    const float x = (float)(10 + (int)floor((double)u01(r[1]) * (double)(wi - 10))), y = (float)(10 + (int)floor((double)u01(r[2]) * (double)(hi - 10)));
    w.d_state[d] = 2;
    w.d_pos[2 * d] = x; w.d_pos[2 * d + 1] = y; w.d_vel[2 * d] = 0.f; w.d_vel[2 * d + 1] = 0.f;
    w.d_ang[d] = 0.f; w.d_angvel[d] = 0.f; w.d_vbias[2 * d] = 0.f; w.d_vbias[2 * d + 1] = 0.f; w.d_wbias[d] = 0.f;
    w.d_mass[d] = 12.f; w.d_mass0[d] = 12.f; w.d_size[d] = 30.f; w.d_elast[d] = 0.5f; w.d_fric[d] = 0.5f; w.d_owner[d] = -1; w.d_eaten[d] = 0.f;
}

// packed poses: (x, y, angle, organism slot) per living cell, (x, y, angle, mass) per dead cell; order is unspecified
__global__ void __launch_bounds__(256) k_pack_frame(WORLD_ARG, FrameWs F, int cell_cap, int dead_cap) {
    const int NC = w.p.n_org * MAXC, NB = NC + w.p.n_dead, lane = lane_id();
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < NB; base += gridDim.x * blockDim.x) {
        int b = base + lane;
        bool cell = false, dead = false;
        if (b < NC) { int o = b >> 6; cell = w.org_state[o] == 1 && (b & 63) < w.org_ncells[o] && w.c_alive[b] == SAPIO_CELL_ALIVE; }
        else if (b < NB) dead = w.d_state[b - NC] != 0;
        uint32_t mc = __ballot_sync(FULL, cell), md = __ballot_sync(FULL, dead);
        int bc = 0, bd = 0;
        if (lane == 0) { if (mc) bc = atomicAdd(&F.counts[0], __popc(mc)); if (md) bd = atomicAdd(&F.counts[1], __popc(md)); }
        bc = __shfl_sync(FULL, bc, 0); bd = __shfl_sync(FULL, bd, 0);
        if (cell) { int i = bc + __popc(mc & ((1u << lane) - 1)); if (i < cell_cap) { float2 p = reinterpret_cast<const float2 *>(w.c_pos)[b]; F.cells[i] = make_float4(p.x, p.y, w.c_ang[b], (float)(b >> 6)); } }
        if (dead) { int d = b - NC, i = bd + __popc(md & ((1u << lane) - 1)); if (i < dead_cap) { float2 p = reinterpret_cast<const float2 *>(w.d_pos)[d]; F.dead[i] = make_float4(p.x, p.y, w.d_ang[d], w.d_mass[d]); } }
    }
}
__global__ void k_frame_header(WORLD_ARG, FrameWs F) {
    SapioFrameHeader h;
    h.co2 = w.g_co2[0]; h.o2 = w.g_o2[0]; h.tick = w.g_tick[0]; h.population = w.g_population[0];
    h.n_cells = F.counts[0]; h.n_dead = F.counts[1];
    for (int i = 0; i < 16; i++) h.stats[i] = w.g_stats[i];
    *F.header = h;
}
