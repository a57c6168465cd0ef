// repro.cuh -- Organism.reproduce / _reproduce_organism (sprites.py:1441-1459), get_new_organism_position (sprites.py:194-219),
// viable_organism_position (sprites.py:156-192), Organism.build_body / build_brain / build_sprite (sprites.py:1683-1729,
// 255-325), and the spawn recipe add_creature_clicked (Sapiogenesis.py:297-323).
//
// Births of one tick, parents in slot order (each sees the newborns accepted before it, as in the reference's list order):
//   k_org_rects        Organism.get_rect of every living organism
//   k_repro_prepare    parallel, one warp per request: DNA.mutate on lane 0, then all 360 bearings tested against the world
//                      bounds and every living organism's rect by the whole warp -> bitmask of viable bearings
//   k_repro_resolve    one warp, sequential over requests: first viable bearing in shuffled order (min Philox key) that also
//                      clears the newborns accepted so far; slot = j-th smallest free slot; rows written (build_sprite)
//   k_repro_finish     parallel, one warp per accepted birth: nn.Linear init of every layer + inherited hidden weights
//                      (mutate_brain_layers carry-over), optimizer state, dead cells inside the footprint deleted
#pragma once
#include "common.cuh"
#include "genome.cuh"
#include "rules.cuh"

__constant__ double C_COS[360], C_SIN[360];      // cos/sin of the integer bearings, filled by the host's libm (== the oracle's)

#define REPRO_NEAR_CAP 64                          // organisms within reach of one birth's candidates; beyond: full scan
struct Birth {
    DGenome g;
    int parent, ok, accepted, slot, brain_changed, old_n; int old_h[16];
    int64_t uid; double px, py, rect[4], dist[2], box[4];
    uint32_t viable[12];
    int near_n, near[REPRO_NEAR_CAP];
    uint8_t order[MAXC];
};
struct ReproWs {
    Birth *births;                 // [birth_cap]
    int *req_flag, *req_off;       // [n_org + 1]
    int *free_flag, *free_off;     // [n_org + 1]
    int *req_list, *free_list;     // [n_org]
    int *accepted;                 // [birth_cap] slots accepted this round
    int *counters;                 // 0 n_req, 1 n_free, 2 n_accepted
    int birth_cap;
};

// viable_organism_position's overlap test between the candidate rect d and a living organism's rect r (sprites.py:180-182)
__device__ __forceinline__ bool rects_block(const double d[4], const double r[4]) {
    bool inx = (d[0] > r[0] && d[0] < r[2]) || (d[2] > r[0] && d[2] < r[2]) || (r[0] > d[0] && r[2] < d[2]);
    if (!inx) return false;
    return (d[1] > r[1] && d[1] < r[3]) || (d[3] > r[1] && d[3] < r[3]) || (r[1] > d[1] && r[3] < d[3]);
}

// Organism.get_rect (sprites.py:1211-1232): seeded with the first cell's centre, widened by living cells' discs
__global__ void __launch_bounds__(256) k_org_rects(WORLD_ARG, const int *n_req) {
    if (n_req && *n_req == 0) return;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < w.p.n_org; o += gridDim.x * blockDim.x) {
        if (w.org_state[o] != 1) continue;
        int r0 = o * MAXC;
        double minx = w.c_pos[2 * r0], maxx = minx, miny = w.c_pos[2 * r0 + 1], maxy = miny;
        for (int k = 0; k < w.org_ncells[o]; k++) {
            int row = r0 + k;
            if (w.c_alive[row] != SAPIO_CELL_ALIVE) continue;
            double x = w.c_pos[2 * row], y = w.c_pos[2 * row + 1], rad = (double)w.c_size[row] / 2.0;
            if (x - rad < minx) minx = x - rad;
            if (x + rad > maxx) maxx = x + rad;
            if (y - rad < miny) miny = y - rad;
            if (y + rad > maxy) maxy = y + rad;
        }
        w.org_rect[4 * o] = minx; w.org_rect[4 * o + 1] = miny; w.org_rect[4 * o + 2] = maxx; w.org_rect[4 * o + 3] = maxy;
    }
}

__global__ void __launch_bounds__(256) k_repro_flags(WORLD_ARG, ReproWs Q) {
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o <= w.p.n_org; o += gridDim.x * blockDim.x) {
        bool in = o < w.p.n_org;
        Q.req_flag[o] = (in && w.org_state[o] == 1 && (w.org_req[o] & SAPIO_REQ_REPRODUCE)) ? 1 : 0;
        Q.free_flag[o] = (in && w.org_state[o] == 0) ? 1 : 0;
    }
}
__global__ void __launch_bounds__(256) k_repro_lists(WORLD_ARG, ReproWs Q) {
    const int n = w.p.n_org;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) {
        if (Q.req_flag[o]) Q.req_list[Q.req_off[o]] = o;
        if (Q.free_flag[o]) Q.free_list[Q.free_off[o]] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { Q.counters[0] = min(Q.req_off[n], Q.birth_cap); Q.counters[1] = Q.free_off[n]; Q.counters[2] = 0;
        if (Q.req_off[n] > Q.birth_cap) { w.g_stats[SAPIO_STAT_OVERFLOW] += 1; w.g_stats[SAPIO_STAT_RARE_PATHS] |= SAPIO_RARE_BIRTH_CAP; } }   // more requests than the tick can serve: counted, never silent
}

#define REPRO_WARPS 2
// phase 1: DNA.mutate (lane 0) + placement candidates against the existing organisms (whole warp)
__global__ void __launch_bounds__(REPRO_WARPS * 32) k_repro_prepare(WORLD_ARG, ReproWs Q, int sibling) {
    const SapioParams &p = w.p;
    const int lane = lane_id(), n_req = Q.counters[0];
    const int32_t T = (int32_t)w.g_tick[0];
    __shared__ DGenome sg[REPRO_WARPS];
    __shared__ GenomeSvc ssvc[REPRO_WARPS];
    for (int r = blockIdx.x * REPRO_WARPS + (threadIdx.x >> 5); r < n_req; r += gridDim.x * REPRO_WARPS) {
        Birth &B = Q.births[r];
        const int parent = Q.req_list[r];
        DGenome &g = sg[threadIdx.x >> 5];                    // the dict-walking logic runs on shared memory, not on global
        GenomeSvc *svc = &ssvc[threadIdx.x >> 5];
        g_from_world(w, parent, g);                           // whole warp
        if (lane != 0) g_svc_worker(g, svc, lane);            // serve lane 0's distance queries until it is done mutating
        else {
            B.parent = parent; B.accepted = 0; B.slot = -1; B.ok = 0;
            double prect[2]; g_get_size(g, prect);                                  // old_organism.dna.get_size()
            RngStream u(p.seed, w.org_uid[parent], T, RNG_GENOME, (uint32_t)sibling << 20);
            int rc = g_mutate_cells(g, p, (double)p.mutation_severity, u, svc);
            g_svc_done(svc);
            bool changed = false;
            if (rc == 0) {
                g_mutate_brain_structure(g, (double)p.brain_mutation_severity, u, B.old_h, B.old_n, changed);
                g_mutate_curiosity(g, (double)p.brain_mutation_severity, u);
                g.generation += 1;
            }
            B.brain_changed = changed;
            if (rc == 0 && g_fits_caps(g, p) && g_instance_order(g, B.order) == g.n) {
                double ns[2]; g_get_size(g, ns);
                B.dist[0] = prect[0] / 1.3 + ns[0] / 1.3; B.dist[1] = prect[1] / 1.3 + ns[1] / 1.3;
                g_get_rect(g, B.rect);
                B.ok = 1;
            }
        }
        __syncwarp();
        {   // publish the child genome for the later phases
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&g); uint32_t *dst = reinterpret_cast<uint32_t *>(&B.g);
            for (int e = lane; e < (int)(sizeof(DGenome) / 4); e += 32) dst[e] = src[e];
        }
        __syncwarp();
        if (lane == 0) {
            B.near_n = 0;
            if (B.ok) {   // reach box of all 360 candidates: organisms outside cannot block any of them
                const double ppx = w.org_pos[2 * parent], ppy = w.org_pos[2 * parent + 1];
                B.box[0] = ppx - B.dist[0] + B.rect[0]; B.box[1] = ppy - B.dist[1] + B.rect[1];
                B.box[2] = ppx + B.dist[0] + B.rect[2]; B.box[3] = ppy + B.dist[1] + B.rect[3];
            }
        }
        __syncwarp();
    }
}

// phase 1b: thread per living organism, all requests from shared-memory tiles: who lies within reach of which birth
#define NEAR_TILE 128
__global__ void __launch_bounds__(256) k_repro_near(WORLD_ARG, ReproWs Q) {
    const int n_req = Q.counters[0];
    if (n_req == 0) return;
    __shared__ double box[NEAR_TILE][4];
    __shared__ uint8_t okv[NEAR_TILE];
    const int n = w.p.n_org;
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const int o = base + threadIdx.x;
        const bool live = o < n && w.org_state[o] == 1;
        double q0 = 0, q1 = 0, q2 = 0, q3 = 0;
        if (live) { q0 = w.org_rect[4 * o]; q1 = w.org_rect[4 * o + 1]; q2 = w.org_rect[4 * o + 2]; q3 = w.org_rect[4 * o + 3]; }
        for (int t0 = 0; t0 < n_req; t0 += NEAR_TILE) {
            __syncthreads();
            for (int e = threadIdx.x; e < NEAR_TILE; e += blockDim.x) {
                int r = t0 + e; bool ok = r < n_req && Q.births[r].ok;
                okv[e] = ok;
                if (ok) { box[e][0] = Q.births[r].box[0]; box[e][1] = Q.births[r].box[1]; box[e][2] = Q.births[r].box[2]; box[e][3] = Q.births[r].box[3]; }
            }
            __syncthreads();
            if (!live) continue;
            const int cnt = min(NEAR_TILE, n_req - t0);
            for (int e = 0; e < cnt; e++) {
                if (!okv[e]) continue;
                if (q2 < box[e][0] || q0 > box[e][2] || q3 < box[e][1] || q1 > box[e][3]) continue;
                int k = atomicAdd(&Q.births[t0 + e].near_n, 1);
                if (k < REPRO_NEAR_CAP) Q.births[t0 + e].near[k] = o;
            }
        }
    }
}

// phase 1c: one warp per request: all 360 bearings against the world bounds and the organisms within reach
__global__ void __launch_bounds__(REPRO_WARPS * 32) k_repro_place(WORLD_ARG, ReproWs Q) {
    const SapioParams &p = w.p;
    const int lane = lane_id(), n_req = Q.counters[0];
    for (int r = blockIdx.x * REPRO_WARPS + (threadIdx.x >> 5); r < n_req; r += gridDim.x * REPRO_WARPS) {
        Birth &B = Q.births[r];
        const int parent = B.parent;
        uint32_t mask[12];
#pragma unroll
        for (int t = 0; t < 12; t++) mask[t] = 0;
        if (B.ok) {
            const double ppx = w.org_pos[2 * parent], ppy = w.org_pos[2 * parent + 1];
            const double r0 = B.rect[0], r1 = B.rect[1], r2 = B.rect[2], r3 = B.rect[3], dx = B.dist[0], dy = B.dist[1];
            // my bearings: b = lane + 32 t; viable so far = inside the world (sprites.py:170-171)
            uint32_t mine = 0;
#pragma unroll
            for (int t = 0; t < 12; t++) {
                int b = lane + 32 * t;
                if (b >= 360) continue;
                double x = ppx + C_COS[b] * dx, y = ppy + C_SIN[b] * dy;
                if (!(r0 + x < 0 || r2 + x > (double)p.width || r1 + y < 0 || r3 + y > (double)p.height)) mine |= 1u << t;
            }
            const int nn = B.near_n;
            const bool listed = nn <= REPRO_NEAR_CAP;
            const int total = listed ? nn : p.n_org;
            if (!listed && lane == 0) { atomicAdd((unsigned long long *)&w.g_stats[SAPIO_STAT_FULL_SCANS], 1ull); atomicOr((unsigned long long *)&w.g_stats[SAPIO_STAT_RARE_PATHS], (unsigned long long)SAPIO_RARE_FULL_SCAN); }
            for (int base = 0; base < total; base += 32) {
                const int e = base + lane;
                int o = -1;
                if (e < total) o = listed ? B.near[e] : e;
                bool near = false;
                double q[4] = {0, 0, 0, 0};
                if (o >= 0 && w.org_state[o] == 1) {
                    q[0] = w.org_rect[4 * o]; q[1] = w.org_rect[4 * o + 1]; q[2] = w.org_rect[4 * o + 2]; q[3] = w.org_rect[4 * o + 3];
                    near = !(q[2] < B.box[0] || q[0] > B.box[2] || q[3] < B.box[1] || q[1] > B.box[3]);
                }
                uint32_t m = __ballot_sync(FULL, near);
                while (m) {
                    int src = __ffs(m) - 1; m &= m - 1;
                    double rr[4];
#pragma unroll
                    for (int c = 0; c < 4; c++) rr[c] = __shfl_sync(FULL, q[c], src);
#pragma unroll
                    for (int t = 0; t < 12; t++) {
                        if (!((mine >> t) & 1u)) continue;
                        int b = lane + 32 * t;
                        double x = ppx + C_COS[b] * dx, y = ppy + C_SIN[b] * dy;
                        double d[4] = {r0 + x, r1 + y, r2 + x, r3 + y};
                        if (rects_block(d, rr)) mine &= ~(1u << t);
                    }
                }
            }
            // transpose per-lane masks into 12 words of 32 bearings
#pragma unroll
            for (int t = 0; t < 12; t++) mask[t] = __ballot_sync(FULL, (mine >> t) & 1u);
        }
        if (lane < 12) {
            uint32_t v = 0;
#pragma unroll
            for (int t = 0; t < 12; t++) if (t == lane) v = mask[t];
            B.viable[lane] = v;
        }
        __syncwarp();
    }
}

// Organism.__init__ + build_body rows (sprites.py:1089-1125, 1698-1729, build_sprite :255-325); whole warp
// g, order and row_of live in shared memory
__device__ __forceinline__ void write_rows(const SapioWorld &w, int slot, const DGenome &g, const uint8_t *order, uint8_t *row_of, double px, double py, int64_t uid, int32_t tick) {
    const int lane = lane_id(), nc = g.n;
    const SapioParams &p_ = w.p;
    float sum_h = 0.f, sum_mh = 0.f, sum_e = 0.f, sum_st = 0.f;
    for (int k = lane; k < nc; k += 32) row_of[order[k]] = (uint8_t)k;
    __syncwarp();
    for (int k = lane; k < MAXC; k += 32) {
        int row = slot * MAXC + k;
        if (k >= nc) { w.c_alive[row] = 0; continue; }
        int j = order[k];
        int prow = -1, mrow = -1;
        if (!g.first[j]) { int pj = g_index_of(g, g.parent_id[j]); if (pj >= 0) prow = row_of[pj]; }
        if (g.mirror_id[j]) { int mj = g_index_of(g, g.mirror_id[j]); if (mj >= 0) mrow = row_of[mj]; }
        w.c_alive[row] = SAPIO_CELL_ALIVE; w.c_type[row] = (uint8_t)g.type[j]; w.c_inuse[row] = 0;
        w.c_parent[row] = (int8_t)prow; w.c_mirror[row] = (int8_t)mrow; w.c_gorder[row] = (uint8_t)j; w.c_id[row] = g.id[j];
        float x = (float)(px + g.relx[j]), y = (float)(py + g.rely[j]);
        reinterpret_cast<float2 *>(w.c_pos)[row] = make_float2(x, y); reinterpret_cast<float2 *>(w.c_spos)[row] = make_float2(x, y);
        reinterpret_cast<float2 *>(w.c_vel)[row] = make_float2(0.f, 0.f); reinterpret_cast<float2 *>(w.c_svel)[row] = make_float2(0.f, 0.f);
        reinterpret_cast<float2 *>(w.c_vbias)[row] = make_float2(0.f, 0.f); reinterpret_cast<float2 *>(w.c_svbias)[row] = make_float2(0.f, 0.f);
        w.c_ang[row] = 0.f; w.c_angvel[row] = 0.f; w.c_wbias[row] = 0.f; w.c_spin_jn[row] = 0.f; w.c_dmg[row] = 0.f;
        float mh = (float)g.maxhealth[j], st = (float)g.storage[j], en = (float)(g.storage[j] / 2.0);
        w.c_health[row] = mh; w.c_energy[row] = en;
        w.c_size[row] = (float)g.size[j]; w.c_mass[row] = (float)g.mass[j]; w.c_elast[row] = (float)g.elast[j]; w.c_fric[row] = (float)g.fric[j];
        reinterpret_cast<float2 *>(w.c_rel)[row] = make_float2((float)g.relx[j], (float)g.rely[j]);
        reinterpret_cast<float2 *>(w.c_grow)[row] = make_float2((float)g.growx[j], (float)g.growy[j]);
        w.c_maxhealth[row] = mh; w.c_use_idle[row] = (float)g.use_idle[j]; w.c_use_act[row] = (float)g.use_act[j]; w.c_storage[row] = st; w.c_damage[row] = (float)g.damage[j];
        sum_h += mh; sum_mh += mh; sum_e += en; sum_st += st;
    }
    sum_h = warp_sum(sum_h); sum_mh = warp_sum(sum_mh); sum_e = warp_sum(sum_e); sum_st = warp_sum(sum_st);
    if (lane == 0) {
        w.org_state[slot] = 1;
        w.org_flags[slot] = (uint8_t)(SAPIO_F_FINITE_MEMORY | (g.mirror_x ? SAPIO_F_MIRROR_X : 0) | (g.mirror_y ? SAPIO_F_MIRROR_Y : 0));
        w.org_ncells[slot] = nc; w.org_uid[slot] = uid; w.org_generation[slot] = g.generation;
        w.org_birth_tick[slot] = w.org_lastbirth_tick[slot] = w.org_think_tick[slot] = w.org_train_tick[slot] = w.org_dopa_tick[slot] = tick;
        w.org_consumed[slot] = 0.f;
        for (int i = 0; i < 4; i++) { w.org_move[4 * slot + i] = 0.f; w.org_out_raw[4 * slot + i] = 0.f; }
        w.org_dopamine[slot] = w.org_pain[slot] = w.org_energy_diff[slot] = 0.f;
        w.org_pos[2 * slot] = (float)px; w.org_pos[2 * slot + 1] = (float)py;
        w.org_curiosity[slot] = (float)g.curiosity; w.org_boredom[slot] = w.org_stimulation[slot] = 0.f;
        for (int i = 0; i < 10; i++) w.org_stim_hist[10 * slot + i] = 0.f;
        w.org_hist_n[slot] = w.org_stm_n[slot] = w.org_mem_n[slot] = 0; w.org_adam_t[slot] = 0; w.org_loss[slot] = 0.f;
        // build_brain (sprites.py:1683-1690): an RNNetwork when use_rnn is set now; lastHidden = zeros (networks.py:149)
        w.org_rnn_h[slot] = p_.rnn_hidden; w.org_hm_n[2 * slot] = w.org_hm_n[2 * slot + 1] = 0; w.org_th_n[slot] = 0; w.org_mem_stamp[slot] = 0;
        for (int i = 0; i < p_.rnn_cap; i++) w.org_last_hidden[(size_t)slot * p_.rnn_cap + i] = 0.f;
        w.org_ic_n[slot] = 0; w.org_req[slot] = 0;
        w.org_gas_a[slot] = 1.0; w.org_gas_b[slot] = 0.0; w.org_gas_in[slot] = 0.0; w.org_gas_refund[slot] = 0.0;
        w.org_last_health[slot] = (sum_mh > 0.f && sum_h != 0.f) ? sum_h / sum_mh * 100.0f : 0.f;      // build_body tail (sprites.py:1728-1729)
        w.org_last_energy[slot] = (sum_st > 0.f && sum_e != 0.f) ? sum_e / sum_st * 100.0f : 0.f;
        // neural.setup_network wiring (neural.py:436-467): eyes then olfactory cells, each in dna.cells order
        int nin = 0;
        for (int pass = 0; pass < 2; pass++)
            for (int j = 0; j < nc; j++) if (g.type[j] == (pass == 0 ? SAPIO_T_EYE : SAPIO_T_OLFACTORY)) w.org_incell[slot * MAXC + nin++] = row_of[j];
        w.org_nincell[slot] = nin;
        int32_t *ldim = w.org_ldim + slot * 16;
        for (int i = 0; i < 16; i++) ldim[i] = 0;
        int nd = 0;
        ldim[nd++] = g_input_size(g);
        for (int i = 0; i < g.n_hidden; i++) ldim[nd++] = g.hidden[i];
        ldim[nd++] = 4;
        w.org_nlayers[slot] = nd - 1;
    }
    __syncwarp();
}

// phase 2: sequential resolve (one warp)
__global__ void __launch_bounds__(32) k_repro_resolve(WORLD_ARG, ReproWs Q, int sibling) {
    const SapioParams &p = w.p;
    const int lane = lane_id(), n_req = Q.counters[0], n_free = Q.counters[1];
    const int32_t T = (int32_t)w.g_tick[0];
    int n_acc = 0;
    for (int r = 0; r < n_req; r++) {
        Birth &B = Q.births[r];
        const int parent = B.parent;
        if (lane == 0) w.org_req[parent] &= (uint8_t)~SAPIO_REQ_REPRODUCE;
        if (!B.ok) { if (lane == 0) stat_add(w, SAPIO_STAT_BIRTH_FAIL_CAPS, 1); continue; }
        // my bearings' shuffle keys: one Philox uniform each, stream position = bearing index
        float key[12]; uint32_t mine = 0;
#pragma unroll
        for (int t = 0; t < 12; t++) {
            int b = lane + 32 * t;
            key[t] = 2.f;
            if (b < 360) { uint32_t rr[4]; rng4(p.seed, w.org_uid[parent], T, RNG_BEARINGS, (uint32_t)(sibling * 90 + b / 4), rr); key[t] = u01(rr[b & 3]); if ((B.viable[t] >> lane) & 1u) mine |= 1u << t; }
        }
        const double ppx = w.org_pos[2 * parent], ppy = w.org_pos[2 * parent + 1];
        bool placed = false, overflow = false;
        for (;;) {
            float bk = 3.f; int bb = 1 << 20;                                   // min (key, bearing): sorted(..., key=random) is stable
#pragma unroll
            for (int t = 0; t < 12; t++) if ((mine >> t) & 1u) { int b = lane + 32 * t; if (key[t] < bk || (key[t] == bk && b < bb)) { bk = key[t]; bb = b; } }
            for (int s = 16; s; s >>= 1) { float k2 = __shfl_xor_sync(FULL, bk, s); int b2 = __shfl_xor_sync(FULL, bb, s); if (k2 < bk || (k2 == bk && b2 < bb)) { bk = k2; bb = b2; } }
            if (bb >= 360) break;
            double x = ppx + C_COS[bb] * B.dist[0], y = ppy + C_SIN[bb] * B.dist[1];
            double d[4] = {B.rect[0] + x, B.rect[1] + y, B.rect[2] + x, B.rect[3] + y};
            bool conflict = false;
            for (int a = lane; a < n_acc; a += 32) { int s = Q.accepted[a]; double rr[4] = {w.org_rect[4 * s], w.org_rect[4 * s + 1], w.org_rect[4 * s + 2], w.org_rect[4 * s + 3]}; conflict |= rects_block(d, rr); }
            if (__any_sync(FULL, conflict)) { if (lane == (bb & 31)) mine &= ~(1u << (bb >> 5)); continue; }
            if (n_acc >= n_free || n_acc >= Q.birth_cap) { overflow = true; break; }
            const int slot = Q.free_list[n_acc];
            const int64_t uid = w.g_next_uid[0];
            // the newborn's Organism.get_rect (sprites.py:1211-1232) from the positions its cells will get: the first cell's
            // centre widened by every cell's disc, all positions as the fp32 the world holds -- for the births still to be placed
            const DGenome &cg = B.g;
            const double hx = (double)(float)x, hy = (double)(float)y;     // first cell: relative_pos (0, 0)
            double minx = hx, maxx = hx, miny = hy, maxy = hy;
            for (int j = lane; j < cg.n; j += 32) {
                double cx = (double)(float)(x + cg.relx[j]), cy = (double)(float)(y + cg.rely[j]), rad = (double)(float)cg.size[j] / 2.0;
                minx = fmin(minx, cx - rad); maxx = fmax(maxx, cx + rad); miny = fmin(miny, cy - rad); maxy = fmax(maxy, cy + rad);
            }
            for (int s = 16; s; s >>= 1) {
                minx = fmin(minx, __shfl_xor_sync(FULL, minx, s)); maxx = fmax(maxx, __shfl_xor_sync(FULL, maxx, s));
                miny = fmin(miny, __shfl_xor_sync(FULL, miny, s)); maxy = fmax(maxy, __shfl_xor_sync(FULL, maxy, s));
            }
            if (lane == 0) {
                w.g_next_uid[0] = uid + 1;
                B.accepted = 1; B.slot = slot; B.uid = uid; B.px = x; B.py = y;
                Q.accepted[n_acc] = slot;
                w.org_rect[4 * slot] = minx; w.org_rect[4 * slot + 1] = miny; w.org_rect[4 * slot + 2] = maxx; w.org_rect[4 * slot + 3] = maxy;
                w.org_state[slot] = 1;                                   // rows follow in k_repro_finish
                stat_add(w, SAPIO_STAT_BIRTHS, 1);
            }
            __syncwarp();
            __threadfence();
            n_acc++;
            placed = true;
            break;
        }
        if (!placed && lane == 0) stat_add(w, overflow ? SAPIO_STAT_OVERFLOW : SAPIO_STAT_BIRTH_FAIL_PLACE, 1);
    }
    if (lane == 0) Q.counters[2] = n_acc;
}

// nn.Linear default init of layer l, element e (W row-major then b): U(-1/sqrt(fan_in), 1/sqrt(fan_in)); one uniform per
// parameter, the stream position IS the parameter's offset in the brain (construction order: input, hidden..., output)
__device__ __forceinline__ float init_param(uint64_t seed, int64_t uid, int32_t tick, uint32_t purpose, int pos, int fan_in) {
    uint32_t r[4];
    rng4(seed, uid, tick, purpose, (uint32_t)(pos >> 2), r);
    double u = (double)u01(r[pos & 3]);
    return (float)((2.0 * u - 1.0) * (1.0 / sqrt((double)fan_in)));
}
__device__ __forceinline__ float stream_u(uint64_t seed, int64_t uid, int32_t tick, uint32_t purpose, uint32_t pos) {
    uint32_t r[4];
    rng4(seed, uid, tick, purpose, pos >> 2, r);
    return u01(r[pos & 3]);
}

// build_brain (sprites.py:1683-1696) for a freshly written slot: fresh init of every layer, then the genome's hidden weights.
// mode 0: spawn -- hidden weights are the tempNet's own init (DNA._setup_brain_structure, sprites.py:868-870)
// mode 1: birth -- hidden weights inherited from `parent` through mutate_brain_layers' carry-over (sprites.py:1022-1044)
__device__ void build_brain(const SapioWorld &w, int slot, int64_t uid, int32_t tick, int mode, int parent, int64_t parent_uid, int sibling,
                            bool brain_changed, const int *old_h, int old_n) {
    const SapioParams &p = w.p;
    const int lane = lane_id();
    const int32_t *ldim = w.org_ldim + slot * 16;
    const int nl = w.org_nlayers[slot];
    float *brain = w.org_brain + (size_t)slot * p.brain_cap;
    // RNN brain (networks.py:111-166): H more columns in the input layer, and the second head outputLayerH constructed after the output layer
    const int H = w.org_rnn_h[slot];
    int off = 0;
    for (int l = 0; l < nl + (H > 0 ? 1 : 0); l++) {
        const int fin = l == nl ? ldim[nl - 1] : ldim[l] + (l == 0 ? H : 0), fout = l == nl ? H : ldim[l + 1], n = fout * fin + fout;
        for (int e = lane; e < n; e += 32) brain[off + e] = init_param(p.seed, uid, tick, RNG_BRAIN_INIT, off + e, fin);
        off += n;
    }
    __syncwarp();
    // hidden layers
    off = ldim[1] * (ldim[0] + H) + ldim[1];
    int poff = 0, fill_base = 0, garb_base = 0;
    const float *pbrain = nullptr; const int32_t *pldim = nullptr;
    if (mode == 1) { pbrain = w.org_brain + (size_t)parent * p.brain_cap; pldim = w.org_ldim + parent * 16; poff = pldim[1] * (pldim[0] + w.org_rnn_h[parent]) + pldim[1]; }
    const int n_old_layers = old_n - 1 > 0 ? old_n - 1 : 0;
    for (int l = 1; l + 1 < nl; l++) {
        const int hl = l - 1, cols = ldim[l], rows = ldim[l + 1];
        if (mode == 0) {
            for (int e = lane; e < rows * cols; e += 32) brain[off + e] = init_param(p.seed, uid, tick, RNG_GENOME_INIT, off + e, cols);
        } else if (!brain_changed) {
            for (int e = lane; e < rows * cols; e += 32) brain[off + e] = pbrain[poff + e];           // same hidden shapes (the input width may differ)
            poff += rows * cols + rows;
        } else {
            const bool have = p.weight_persistence && hl < n_old_layers;
            const int oc = have ? old_h[hl] : 0, orows = have ? old_h[hl + 1] : 0, old_len = oc * orows;
            const int rhave = have ? min(rows, orows) : 0;
            for (int r = lane; r < rows; r += 32) {
                if (r < rhave) {
                    // row view resize_(): re-reads the old matrix storage from the row's offset; the tail past the storage end
                    // is uninitialised in the reference -> U(0,1) from its own stream (see oracle/genome.c)
                    int gb = garb_base;
                    for (int r2 = 0; r2 < r; r2++) { int cs = max(0, old_len - r2 * oc); gb += max(0, cols - cs); }
                    for (int c = 0; c < cols; c++) {
                        int idx = r * oc + c;
                        brain[off + r * cols + c] = idx < old_len ? pbrain[poff + idx] : stream_u(p.seed, parent_uid, tick, RNG_WEIGHT_GARBAGE, ((uint32_t)sibling << 20) + (uint32_t)(gb++));
                    }
                } else {
                    int fb = fill_base + (r - rhave) * cols;
                    for (int c = 0; c < cols; c++) brain[off + r * cols + c] = stream_u(p.seed, parent_uid, tick, RNG_WEIGHT_FILL, ((uint32_t)sibling << 20) + (uint32_t)(fb + c));
                }
            }
            fill_base += (rows - rhave) * cols;
            for (int r2 = 0; r2 < rhave; r2++) { int cs = max(0, old_len - r2 * oc); garb_base += max(0, cols - cs); }
            if (have) poff += old_len + orows;        // next old hidden layer in the parent's brain (skip this one's W and b)
        }
        off += rows * cols + rows;
    }
    // what DNA.mutate's deepcopy carries over and erase_memory (sprites.py:413-421) does not clear: hidden_memory (its first
    // SAPIO_HM - SAPIO_STM entries: the list never exceeds SAPIO_HM here) and trainingHidden -- see oracle/world.c ow_inherit_recurrent
    if (mode == 1 && p.rnn_cap > 0) {
        const int RC = p.rnn_cap, app = w.org_hm_n[2 * parent], pop = w.org_hm_n[2 * parent + 1], n = min(app - pop, SAPIO_HM - SAPIO_STM), tn = w.org_th_n[parent];
        const float *ph = w.org_hm + (size_t)parent * SAPIO_HM * RC; float *ch = w.org_hm + (size_t)slot * SAPIO_HM * RC;
        for (int e = lane; e < n * RC; e += 32) { const int i = e / RC, c = e - i * RC; ch[e] = ph[(size_t)((pop + i) % SAPIO_HM) * RC + c]; }
        const float *pt = w.org_th + (size_t)parent * SAPIO_MEMROWS * RC; float *ct = w.org_th + (size_t)slot * SAPIO_MEMROWS * RC;
        for (int e = lane; e < tn * RC; e += 32) ct[e] = pt[e];
        if (lane == 0) { w.org_hm_n[2 * slot] = n; w.org_hm_n[2 * slot + 1] = 0; w.org_th_n[slot] = tn; }
    }
    // optimizer state, joint impulses
    const size_t ac = (size_t)p.adam_cap;
    for (size_t e = lane; e < ac; e += 32) { w.org_adam_m[(size_t)slot * ac + e] = 0.f; w.org_adam_v[(size_t)slot * ac + e] = 0.f; }
    for (int e = lane; e < SAPIO_NPAIR; e += 32) w.j_jn[(size_t)slot * SAPIO_NPAIR + e] = 0.f;
    __syncwarp();
}

// phase 3: brains + footprint (dead cells whose centre lies strictly inside the newborn's rect are deleted, sprites.py:184-217)
__global__ void __launch_bounds__(REPRO_WARPS * 32) k_repro_finish(WORLD_ARG, ReproWs Q, int sibling) {
    const int lane = lane_id(), n_req = Q.counters[0];
    const int32_t T = (int32_t)w.g_tick[0];
    __shared__ DGenome sg[REPRO_WARPS];
    __shared__ uint8_t sorder[REPRO_WARPS][MAXC], srow_of[REPRO_WARPS][MAXC];
    for (int r = blockIdx.x * REPRO_WARPS + (threadIdx.x >> 5); r < n_req; r += gridDim.x * REPRO_WARPS) {
        Birth &B = Q.births[r];
        if (!B.accepted) continue;
        DGenome &g = sg[threadIdx.x >> 5];
        {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&B.g); uint32_t *dst = reinterpret_cast<uint32_t *>(&g);
            for (int e = lane; e < (int)(sizeof(DGenome) / 4); e += 32) dst[e] = src[e];
            for (int k = lane; k < MAXC; k += 32) sorder[threadIdx.x >> 5][k] = B.order[k];
        }
        __syncwarp();
        write_rows(w, B.slot, g, sorder[threadIdx.x >> 5], srow_of[threadIdx.x >> 5], B.px, B.py, B.uid, T);
        build_brain(w, B.slot, B.uid, T, 1, B.parent, w.org_uid[B.parent], sibling, B.brain_changed != 0, B.old_h, B.old_n);
    }
}
// build_body's footprint clearing (sprites.py:1707-1713): a dead cell whose centre lies strictly inside the newborn's DNA rect
// is deleted. Thread per dead cell against the accepted births of this round (shared-memory tiles).
__global__ void __launch_bounds__(256) k_repro_clear(WORLD_ARG, ReproWs Q) {
    const int n_req = Q.counters[0];
    if (Q.counters[2] == 0) return;
    __shared__ double fp[NEAR_TILE][4];
    __shared__ uint8_t acc[NEAR_TILE];
    __shared__ int s_removed;
    if (threadIdx.x == 0) s_removed = 0;
    const int n = w.p.n_dead;
    int removed = 0;
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const int d = base + threadIdx.x;
        const bool live = d < n && w.d_state[d] != 0;
        double x = 0, y = 0;
        if (live) { x = w.d_pos[2 * d]; y = w.d_pos[2 * d + 1]; }
        bool hit = false;
        for (int t0 = 0; t0 < n_req; t0 += NEAR_TILE) {
            __syncthreads();
            for (int e = threadIdx.x; e < NEAR_TILE; e += blockDim.x) {
                int r = t0 + e; bool a = r < n_req && Q.births[r].accepted;
                acc[e] = a;
                if (a) { const Birth &B = Q.births[r]; fp[e][0] = B.rect[0] + B.px; fp[e][1] = B.rect[1] + B.py; fp[e][2] = B.rect[2] + B.px; fp[e][3] = B.rect[3] + B.py; }
            }
            __syncthreads();
            if (!live || hit) continue;
            const int cnt = min(NEAR_TILE, n_req - t0);
            for (int e = 0; e < cnt && !hit; e++) hit = acc[e] && x > fp[e][0] && x < fp[e][2] && y > fp[e][1] && y < fp[e][3];
        }
        if (hit) { w.d_state[d] = 0; removed++; }
    }
    if (removed) atomicAdd(&s_removed, removed);
    __syncthreads();
    if (threadIdx.x == 0 && s_removed) stat_add(w, SAPIO_STAT_DEAD_REMOVED, s_removed);
}
// siblings: a parent whose offspring i was not born stops reproducing this tick (sprites.py:1456-1459)
__global__ void k_repro_next_sibling(WORLD_ARG, ReproWs Q) {
    const int n_req = Q.counters[0];
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_req; r += gridDim.x * blockDim.x)
        if (Q.births[r].accepted) w.org_req[Q.births[r].parent] |= SAPIO_REQ_REPRODUCE;
}

// ---- spawn: add_creature_clicked (Sapiogenesis.py:297-323) = DNA().randomize(widget ranges) + Organism + build_body ----
__global__ void k_spawn_finish(WORLD_ARG, int n) { w.g_next_uid[0] += n; }
__global__ void __launch_bounds__(REPRO_WARPS * 32) k_spawn(WORLD_ARG, const int *slots, const double *pos, int n) {
    const SapioParams &p = w.p;
    const int lane = lane_id();
    const int32_t T = (int32_t)w.g_tick[0];
    const int64_t uid0 = w.g_next_uid[0];
    __shared__ DGenome sg[REPRO_WARPS];
    __shared__ uint8_t sorder[REPRO_WARPS][MAXC];
    __shared__ int sok[REPRO_WARPS];
    __shared__ uint8_t srow_of[REPRO_WARPS][MAXC];
    DGenome &g = sg[threadIdx.x >> 5];
    uint8_t *order = sorder[threadIdx.x >> 5];
    for (int i = blockIdx.x * REPRO_WARPS + (threadIdx.x >> 5); i < n; i += gridDim.x * REPRO_WARPS) {
        const int64_t uid = uid0 + i;
        if (lane == 0) {
            int ok = 0;
            for (int attempt = 0; attempt < 64 && !ok; attempt++) {              // re-roll genomes that exceed the tile caps (ours)
                RngStream u(p.seed, uid, T, RNG_GENOME, (uint32_t)attempt << 18);
                if (g_randomize(g, p, u) < 0 || !g_fits_caps(g, p)) continue;
                ok = g_instance_order(g, order) == g.n;
            }
            sok[threadIdx.x >> 5] = ok;
        }
        __syncwarp();
        if (sok[threadIdx.x >> 5]) {
            write_rows(w, slots[i], g, order, srow_of[threadIdx.x >> 5], pos[2 * i], pos[2 * i + 1], uid, T);
            build_brain(w, slots[i], uid, T, 0, -1, 0, 0, false, nullptr, 0);
        } else if (lane == 0) stat_add(w, SAPIO_STAT_BIRTH_FAIL_CAPS, 1);
        __syncwarp();
    }
}
