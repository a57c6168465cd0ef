// think.cuh -- neural.activate (neural.py:91-366) for every organism whose brain is due: one warp per organism.
//   sensors  : each eye / olfactory cell's `view` is the set of solid bodies of other organisms whose disc overlaps the
//              sensor disc (centre: the sensor BODY, sprites.py:300-323); members are found through the broadphase grid,
//              lanes striding over the grid cells the disc covers, exact double predicate (contract C3)
//   encoding : 8-sector eye (neural.py:123-196), olfactory (neural.py:197-218), speed/x/y, energy (neural.py:227-244)
//   forward  : networks.Network.forward (networks.py:301-308), warp-level FMA over the ragged layer chain (batch of one
//              per organism: a GEMV chain, bound by reading the weights once)
//   after    : 3-decimal rounding, boredom gate + randn (Philox), history / short-term memory rings, movement clamp,
//              calculateStimulation + boredom (neural.py:60-75, 351-366)
#pragma once
#include "common.cuh"
#include "rng.cuh"
#include "rules.cuh"

__device__ __forceinline__ double warp_max(double v) { for (int s = 16; s; s >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, s)); return v; }
__device__ __forceinline__ double warp_sum_d(double v) { for (int s = 16; s; s >>= 1) v += __shfl_xor_sync(FULL, v, s); return v; }

struct SenseOut { double eye[8]; double smell; int hits; };

// The view of the sensor disc at (sx, sy) radius R of organism o: lanes stride over the grid cells the disc can reach and call
// on_hit(body id, x, y, r, dead) for every member (exact double predicate, contract C3). The one enumeration of the library:
// k_think encodes from it, the parity probe k_sensor_view lists its hits.
template <class F>
__device__ __forceinline__ void sense_enumerate(const SapioWorld &w, const Workspace &W, int o, int64_t uid, float sx, float sy, float R, F &&on_hit) {
    const SapioParams &p = w.p;
    const int NC = p.n_org * MAXC, lane = lane_id();
    const float gc = p.grid_cell, reach = R + 0.5f * gc;            // grid_cell >= 2 * largest solid radius
    int cx0 = (int)floorf((sx - reach - p.grid_x0) / gc) + 1, cx1 = (int)floorf((sx + reach - p.grid_x0) / gc) + 1;
    int cy0 = (int)floorf((sy - reach) / gc) + 1, cy1 = (int)floorf((sy + reach) / gc) + 1;
    cx0 = max(cx0, 0); cy0 = max(cy0, 0); cx1 = min(cx1, p.grid_nx - 1); cy1 = min(cy1, p.grid_ny - 1);
    const int nbx = cx1 - cx0 + 1, nby = cy1 - cy0 + 1;
    if (nbx <= 0 || nby <= 0) return;
    for (int idx = lane; idx < nbx * nby; idx += 32) {
        int c = (cy0 + idx / nbx) * p.grid_nx + cx0 + idx % nbx;
        for (int i = W.cell_start[c], e = W.cell_start[c + 1]; i < e; i++) {
            // the grid's own record of the body -- position, radius, id as written by k_grid_rank from the same arrays (positions only change
            // when the grid is rebuilt; every listed body is a living cell or an existing dead cell): one 16-byte load instead of the chain
            // body id -> c_alive -> c_pos -> c_size of dependent cold loads
            const float4 rec = W.srec[i];
            const int j = __float_as_int(rec.w);
            const float x = rec.x, y = rec.y, r = rec.z;
            bool dead = false;
            if (j < NC) {
                if ((j >> 6) == o) continue;
            } else {
                if (w.d_owner[j - NC] == uid) continue;                        // own dead cells are invisible (sprites.py:233)
                dead = true;
            }
            double d2;
            if (!circles_overlap(sx, sy, R, x, y, r, &d2)) continue;
            on_hit(j, x, y, r, dead);
        }
    }
}

// Encodes the view relative to the cell at (px, py) (neural.py:123-218).
template <bool EYE>
__device__ __forceinline__ void sense(const SapioWorld &w, const Workspace &W, int o, int64_t uid, float sx, float sy, float R, float px, float py, double rot_c, double rot_s, SenseOut &out) {
    double eye[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
    double smell = 0.; int hits = 0;
    sense_enumerate(w, W, o, uid, sx, sy, R, [&](int, float x, float y, float r, bool dead) {
        hits++;
        // encoding in double like the reference's Python floats: the values are rounded to 3 decimals next, and a
        // 1e-7 fp32 error would flip a digit at every tie; this path runs once per think, not per tick
        double ddx = (double)x - (double)px, ddy = (double)y - (double)py;
        double dist = sqrt(ddx * ddx + ddy * ddy);
        double val = fabs(((double)R - (dist + (double)r)) / (double)R);
        if (EYE) {
            // (cos, sin) of atan2(ddy, ddx) - rotation: the direction of the hit turned back by the organism's rotation -- a dot and a cross
            // product with the rotation's unit vector over the distance already at hand, instead of three double-precision
            // transcendentals per hit (same value to the last bits of a double; atan2(0, 0) = 0)
            const double ux = dist > 0.0 ? ddx / dist : 1.0, uy = dist > 0.0 ? ddy / dist : 0.0;
            const double dx = ux * rot_c + uy * rot_s, dy = uy * rot_c - ux * rot_s;
            int s;
            if (dx >= 0.0) { if (dy < 0.0) s = dx < .707 ? 0 : 1; else s = dx >= .707 ? 2 : 3; }
            else { if (dy >= 0.0) s = dy >= .707 ? 4 : 5; else s = dx < -.707 ? 6 : 7; }
#pragma unroll
            for (int q = 0; q < 8; q++) if (q == s) eye[q] = fmax(eye[q], val);
        } else if (dead) smell = fmax(smell, val);
    });
    if (EYE) {
#pragma unroll
        for (int q = 0; q < 8; q++) out.eye[q] = warp_max(eye[q]);
    } else out.smell = warp_max(smell);
    for (int s = 16; s; s >>= 1) hits += __shfl_xor_sync(FULL, hits, s);
    out.hits = hits;
}

#define THINK_WARPS 4
#ifndef THINK_MINB
#define THINK_MINB 4          // resident CTAs per SM the kernel is compiled for. Measured (A/B builds, 100k organisms aged 1500 ticks): 8 CTAs at <= 64
#endif                        // registers 0.55 ms, 4 CTAs at 117 registers 0.52 ms -- the organisms in flight are not what bounds the stage
// dynamic shared memory per warp: in_cap + 2 * max(width_cap, SAPIO_RNN_MAXH) floats
__global__ void __launch_bounds__(THINK_WARPS * 32, THINK_MINB) k_think(WORLD_ARG, Workspace W, RulesWs R) {
    extern __shared__ float think_smem[];
    const SapioParams &p = w.p;
    const int IN = p.in_cap, lane = lane_id();
    float *x = think_smem + (size_t)(threadIdx.x >> 5) * (IN + 2 * max(p.width_cap, SAPIO_RNN_MAXH));
    float *act0 = x + IN, *act1 = act0 + max(p.width_cap, SAPIO_RNN_MAXH);
    const int n_think = R.counters[RC_NTHINK];
    const int32_t T = (int32_t)w.g_tick[0];
    for (;;) {                                          // organisms differ by their sensors and brains: pulled from a work counter, not dealt
        int q = 0;
        if (lane == 0) q = atomicAdd(&R.counters[RC_THINK_TICKET], 1);
        q = __shfl_sync(FULL, q, 0);
        if (q >= n_think) break;
        const int o = R.think_list[q];
        const int nc = w.org_ncells[o], nl = w.org_nlayers[o];
        const int32_t *ldim = w.org_ldim + o * 16;
        const int I = ldim[0];
        const int H = w.org_rnn_h[o], RC = p.rnn_cap;                                  // RNN brain: H hidden values follow the I inputs (I + H <= in_cap: g_fits_caps)
        const int64_t uid = w.org_uid[o];
        // Organism.rotation (sprites.py:1179-1201) as an angle
        uint32_t a0 = __ballot_sync(FULL, lane < nc && w.c_alive[o * MAXC + lane] == SAPIO_CELL_ALIVE);
        uint32_t a1 = __ballot_sync(FULL, lane + 32 < nc && w.c_alive[o * MAXC + lane + 32] == SAPIO_CELL_ALIVE);
        uint64_t alive = ((uint64_t)a1 << 32) | a0;
        double rot_c = 1.0, rot_s = 0.0;                                               // cos, sin of the rotation: angle(heart -> first other living cell) now minus in the genome's frame
        {
            uint64_t am = alive & ~1ull;
            if (am) {
                int rel = __ffsll((long long)am) - 1, r0 = o * MAXC, rr = r0 + rel;
                const double rx = (double)w.c_pos[2 * rr] - (double)w.c_pos[2 * r0], ry = (double)w.c_pos[2 * rr + 1] - (double)w.c_pos[2 * r0 + 1];
                const double ox = (double)w.c_rel[2 * rr], oy = (double)w.c_rel[2 * rr + 1];
                const double nr = sqrt(rx * rx + ry * ry), no = sqrt(ox * ox + oy * oy);
                const double cx = nr > 0.0 ? rx / nr : 1.0, cy = nr > 0.0 ? ry / nr : 0.0, qx = no > 0.0 ? ox / no : 1.0, qy = no > 0.0 ? oy / no : 0.0;
                rot_c = cx * qx + cy * qy; rot_s = cy * qx - cx * qy;
            }
        }
        int n = 0;
        for (int qi = 0; qi < w.org_nincell[o]; qi++) {
            int k = w.org_incell[o * MAXC + qi], row = o * MAXC + k, t = w.c_type[row];
            bool live = (alive >> k) & 1ull;
            float2 cp = reinterpret_cast<const float2 *>(w.c_pos)[row], sp = reinterpret_cast<const float2 *>(w.c_spos)[row];
            SenseOut so;
            if (t == SAPIO_T_EYE) {
#pragma unroll
                for (int s = 0; s < 8; s++) so.eye[s] = 0.;
                if (live) sense<true>(w, W, o, uid, sp.x, sp.y, w.c_size[row] * p.eyesight_mult, cp.x, cp.y, rot_c, rot_s, so);
                if (lane < 8) { double v = 0.;
#pragma unroll
                    for (int s = 0; s < 8; s++) if (s == lane) v = so.eye[s];
                    x[n + lane] = (float)v; }
                n += 8;
            } else {
                so.smell = 0.;
                if (live) sense<false>(w, W, o, uid, sp.x, sp.y, w.c_size[row] * p.olfactory_mult, cp.x, cp.y, rot_c, rot_s, so);
                if (lane == 0) x[n] = (float)so.smell;
                n += 1;
            }
        }
        {   // speed, x_direction, y_direction (last commanded), energy fraction (neural.py:227-244)
            double cur = 0., mx = 0.;
            for (int k = lane; k < nc; k += 32) { int row = o * MAXC + k; if ((alive >> k) & 1ull) cur += (double)w.c_energy[row]; mx += (double)w.c_storage[row]; }
            cur = warp_sum_d(cur); mx = warp_sum_d(mx);
            if (lane == 0) { x[n] = w.org_move[4 * o + 1]; x[n + 1] = w.org_move[4 * o + 2]; x[n + 2] = w.org_move[4 * o + 3]; x[n + 3] = (float)((cur / mx * 100.0) / 100.); }
            n += 4;
        }
        __syncwarp();
        for (int i = lane; i < I; i += 32) x[i] = rintf(x[i] * 1000.0f) / 1000.0f;                 // neural.py:285 (round half to even)
        if (lane < H) x[I + lane] = w.org_last_hidden[(size_t)o * RC + lane];                      // torch.cat((inputs, hidden)) (networks.py:164), neural.py:287-288
        __syncwarp();
        // Network.forward: sigmoid after the first layer only. RNNetwork.forward (networks.py:160-175): one more "layer" after the
        // output layer -- the head outputLayerH on the same last hidden activation -- gives the next lastHidden
        const float *brain = w.org_brain + (size_t)o * p.brain_cap;
        const float *cur = x; float *dst = act0;
        float out4[4] = {0.f, 0.f, 0.f, 0.f};
        int off = 0;
        for (int l = 0; l < nl + (H > 0 ? 1 : 0); l++) {
            const bool head = l == nl;
            const int fin = head ? ldim[nl - 1] : ldim[l] + (l == 0 ? H : 0), fout = head ? H : ldim[l + 1];
            const float *Wm = brain + off, *b = Wm + fout * fin;
            // eight output rows at a time: their weights (cold, straight from HBM) are all requested before the first is used.
            // Per row the summation order is unchanged: lane-strided partial sums in ascending column order, then the butterfly.
            for (int r0 = 0; r0 < fout; r0 += 8) {
                float acc[8], wv[8][4], bv[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int r = r0 + u;
                    bv[u] = r < fout ? b[r] : 0.f;
#pragma unroll
                    for (int cc = 0; cc < 4; cc++) { const int c = lane + 32 * cc; wv[u][cc] = (r < fout && c < fin) ? Wm[r * fin + c] : 0.f; }
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    float sacc = 0.f;
#pragma unroll
                    for (int cc = 0; cc < 4; cc++) { const int c = lane + 32 * cc; if (c < fin) sacc += wv[u][cc] * cur[c]; }
                    acc[u] = sacc;
                }
#pragma unroll
                for (int sh = 16; sh; sh >>= 1) {
#pragma unroll
                    for (int u = 0; u < 8; u++) acc[u] += __shfl_xor_sync(FULL, acc[u], sh);
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int r = r0 + u;
                    if (r >= fout) break;
                    float sv = acc[u] + bv[u];
                    if (l == 0) sv = 1.0f / (1.0f + expf(-sv));
                    if (l == nl - 1) {
#pragma unroll
                        for (int z = 0; z < 4; z++) if (z == r) out4[z] = sv;
                    } else if (lane == 0) dst[r] = sv;
                }
            }
            __syncwarp();
            if (l < nl - 1) { cur = dst; dst = (dst == act0) ? act1 : act0; }                   // the output layer wrote registers: `dst` stays free for the head, `cur` stays its input
            off += fout * fin + fout;
        }
        const float *hid = dst;                                                                    // H values when H > 0
        if (lane < H) w.org_last_hidden[(size_t)o * RC + lane] = hid[lane];
        if (lane < 4) { float v = 0.f;
#pragma unroll
            for (int z = 0; z < 4; z++) if (z == lane) v = out4[z];
            w.org_out_raw[4 * o + lane] = v; }
        float outv[4];
#pragma unroll
        for (int z = 0; z < 4; z++) outv[z] = rintf(out4[z] * 1000.0f) / 1000.0f;                  // neural.py:292
        uint32_t rr4[4];
        rng4(p.seed, uid, T, RNG_BOREDOM, 0, rr4);
        if (u01(rr4[0]) <= w.org_boredom[o]) {                                                     // neural.py:294-295
            rng4(p.seed, uid, T, RNG_RANDN, 0, rr4);
#pragma unroll
            for (int h = 0; h < 2; h++) {                                                          // Box-Muller on (0,1] x [0,1)
                // in double, rounded once: the action drives the push actuator, and a push one ulp off flips position roundings (rules.cuh)
                const double u1 = ((double)(rr4[2 * h] >> 8) + 1.0) / 16777216.0, u2 = (double)u01(rr4[2 * h + 1]);
                const double rad = sqrt(-2.0 * log(u1)), th = 2.0 * 3.14159265358979323846 * u2;
                outv[2 * h] = (float)(rad * cos(th)); outv[2 * h + 1] = (float)(rad * sin(th));
            }
        }
        // history rings (slot = running count % capacity), short-term memory (neural.py:301-318)
        const int cnt = w.org_hist_n[o], nv = min(cnt, SAPIO_HIST);
        float *hin = w.org_hist_in + (size_t)o * SAPIO_HIST * IN, *hout = w.org_hist_out + o * 40, *hdo = w.org_hist_dopa + o * 10;
        const float dopamine = w.org_dopamine[o], pain = w.org_pain[o];
        if ((fabsf(dopamine) > 0.1f || fabsf(pain) > 0.1f) && nv + 1 > 3) {
            const int s4 = (cnt - 3) % SAPIO_HIST, sc = w.org_stm_n[o], slot = sc % SAPIO_STM;
            float *sin_ = w.org_stm_in + ((size_t)o * SAPIO_STM + slot) * IN;
            for (int i = lane; i < I; i += 32) sin_[i] = hin[(size_t)s4 * IN + i];
            if (lane < 4) w.org_stm_out[(o * SAPIO_STM + slot) * 4 + lane] = hout[s4 * 4 + lane];
            if (lane == 0) {
                double d2 = hdo[(cnt - 2) % SAPIO_HIST], d1 = hdo[(cnt - 1) % SAPIO_HIST];
                w.org_stm_rew[o * SAPIO_STM + slot] = (float)((((0.0 + d2) + d1) + (double)dopamine) / 3);
                w.org_stm_n[o] = sc + 1;
            }
            if (H > 0) {                                                                           // hidden_memory: append, pop with the short-term memory (neural.py:312-318)
                const int app = w.org_hm_n[2 * o];
                if (lane < H) w.org_hm[((size_t)o * SAPIO_HM + (size_t)(app % SAPIO_HM)) * RC + lane] = hid[lane];
                if (lane == 0) { w.org_hm_n[2 * o] = app + 1; if (sc + 1 > SAPIO_STM) w.org_hm_n[2 * o + 1]++; }
            }
        }
        __syncwarp();
        const int slot = cnt % SAPIO_HIST;
        for (int i = lane; i < I; i += 32) hin[(size_t)slot * IN + i] = x[i];
        if (lane < 4) {
            float v = 0.f;
#pragma unroll
            for (int z = 0; z < 4; z++) if (z == lane) v = outv[z];
            hout[slot * 4 + lane] = v;
            w.org_move[4 * o + lane] = fminf(fmaxf(v, -1.0f), 1.0f);                               // neural.py:320-349
        }
        if (lane == 0) { hdo[slot] = dopamine; w.org_hist_n[o] = cnt + 1; }
        __syncwarp();
        // calculateStimulation (neural.py:60-75): fp32, oldest entry first
        const int nn = min(cnt + 1, SAPIO_HIST);
        float part = 0.f;
        for (int i = lane; i < I; i += 32) {
            float s = 0.f;
            for (int e = 0; e < nn; e++) { int sl = (cnt + 1 - nn + e) % SAPIO_HIST; s = s + (sl == slot ? x[i] : hin[(size_t)sl * IN + i]); }
            part += fabsf(s / (float)nn);
        }
        const float new_stim = warp_sum(part) / (float)I;
        if (lane == 0) {
            float *sh = w.org_stim_hist + 10 * o;                                                  // neural.py:354-366
            double avg = 0.0;
            for (int i = 0; i < 9; i++) { float v = sh[i + 1]; sh[i] = v; avg += (double)v; }
            sh[9] = new_stim; avg = (avg + (double)new_stim) / 10;
            double stimulation = fabs((double)new_stim - avg), cu = w.org_curiosity[o];
            w.org_stimulation[o] = (float)stimulation;
            w.org_boredom[o] = (float)((stimulation != 0.0 && cu != 0.0) ? (1.0 - stimulation) * cu : 0.0);
        }
        __syncwarp();
    }
}

// parity probe: the view of one sensor cell as a sorted id list ("sensor hit indices bit-exact"), listed by the production
// enumeration above (one warp, like k_think)
__global__ void k_sensor_view(WORLD_ARG, Workspace W, int o, int k, uint32_t *out, int cap, int *out_n) {
    const SapioParams &p = w.p;
    const int row = o * MAXC + k;
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    float sx = w.c_spos[2 * row], sy = w.c_spos[2 * row + 1];
    float R = w.c_size[row] * (w.c_type[row] == SAPIO_T_EYE ? p.eyesight_mult : p.olfactory_mult);
    if (threadIdx.x == 0) *out_n = 0;
    __syncwarp();
    sense_enumerate(w, W, o, w.org_uid[o], sx, sy, R, [&](int j, float, float, float, bool) {
        int n = atomicAdd(out_n, 1);
        if (n < cap) out[n] = (uint32_t)j;
    });
    __syncwarp();
    if (threadIdx.x == 0) {
        const int n = min(*out_n, cap);
        for (int i = 1; i < n; i++) { uint32_t v = out[i]; int j = i - 1; while (j >= 0 && out[j] > v) { out[j + 1] = out[j]; j--; } out[j + 1] = v; }
    }
}
