// rules.cuh -- sprites.update_organisms (sprites.py:1809-1856) on the GPU: one warp per organism, two cell rows per lane.
//
//   k_rules_begin   age limit, digestion, think gate (+ _update_brain dopamine), and the organism's gas map:
//                   which cells reach their turn alive is decided by start-of-tick health and the growth tree alone,
//                   so each organism's effect on the global (co2, o2) pair is an affine map known before the main pass
//   inclusive scan  ordered composition of those maps (they do not commute): every organism learns the co2 level it
//                   would have met in the reference's sequential loop (contract C4, SURVEY C-6)
//   k_think         neural.activate for organisms whose brain is due                                (think.cuh)
//   k_rules_main    the rest of Organism.update: snapshots, heart-only death, reproduction gate, the per-cell loop
//                   (_update_cell: death cascade, carnivory, push / rotate actuators, energy use, gas exchange, heal, starve)
//   k_rules_finish  global gases after the phase (scan total + deferred refunds)
//   settle kernels  deferred carnivore damage / eaten mass (contract C2), dying cells -> dead pool in deterministic slot
//                   order, gradual_dispersion, bounds, organism slots freed
#pragma once
#include <cub/cub.cuh>
#include "common.cuh"
#include "rng.cuh"
#include "step.cuh"

struct Aff { double a, b; };                        // x -> a*x + b on the co2 level
struct AffCompose { __host__ __device__ Aff operator()(const Aff &first, const Aff &then) const { return Aff{then.a * first.a, then.a * first.b + then.b}; } };

struct RulesWs {
    Aff *gas_map, *gas_scan;        // [n_org]
    int *think_list;                // [n_org]
    int *counters;                  // [16]: 0 n_think, 1 n_train, 2 n_repro, 3 n_dying, 4 n_free
    double *refund_sum;             // [1]
    int *dying_cnt, *dying_off;     // [n_org + 1]
    int *free_flag, *free_off;      // [n_dead + 1]
    int *free_list;                 // [n_dead]
    double *disp_factor;            // [n_dead] (1 - r) of each dispersing dead cell
    double *disp_prod;              // [1]
    int *train_list, *repro_list;   // [n_org]
    // slab mode (p.slab_mode, sapiogenesis_b200/slabs.py): the co2 level the first organism of this slab meets (the slabs before it composed by
    // the host), and which dead slots hold copies of a neighbour's dead cells (no part in the dispersion's gas exchange)
    double *slab_co2_in;            // [1]
    uint8_t *dead_ghost;            // [n_dead]
};
enum { RC_NTHINK = 0, RC_NTRAIN = 1, RC_NREPRO = 2, RC_NDYING = 3, RC_NFREE = 4, RC_TRAIN_TICKET = 8 /* work counter of k_train */, RC_THINK_TICKET = 9 /* work counter of k_think */ };

__device__ __forceinline__ float warp_sum(float v) { for (int s = 16; s; s >>= 1) v += __shfl_xor_sync(FULL, v, s); return v; }
__device__ __forceinline__ float bcast2(float v0, float v1, int k) { return __shfl_sync(FULL, (k >> 5) ? v1 : v0, k & 31); }
__device__ __forceinline__ int bcast2i(int v0, int v1, int k) { return __shfl_sync(FULL, (k >> 5) ? v1 : v0, k & 31); }

// per-lane view of one organism: cell rows lane and lane + 32
struct OrgR {
    int o, nc, lane;
    float health[2], storage[2], maxhealth[2], size[2], use_idle[2], use_act[2], damage[2];
    double energy[2];                // Python floats in the reference: an organism's energies pass through ~2n dependent updates per tick and the co2
                                     // refund is what is left of them; held in double here, rounded once when stored
    float vx[2], vy[2], av[2], px[2], py[2];
    int alive[2], type[2], inuse[2], parent[2];
    uint64_t anc[2];                 // ancestors in the growth tree (dna.grows_from chain)

    __device__ void load(const SapioWorld &w, int o_, bool with_body) {
        o = o_; nc = w.org_ncells[o]; lane = lane_id();
#pragma unroll
        for (int h = 0; h < 2; h++) {
            int k = lane + 32 * h, row = o * MAXC + k;
            bool in = k < nc;
            alive[h] = in ? w.c_alive[row] : 0;
            type[h] = in ? w.c_type[row] : 0;
            parent[h] = in ? (int)w.c_parent[row] : -1;
            health[h] = in ? w.c_health[row] : 0.f; energy[h] = in ? (double)w.c_energy[row] : 0.0;
            storage[h] = in ? w.c_storage[row] : 0.f; maxhealth[h] = in ? w.c_maxhealth[row] : 0.f;
            size[h] = in ? w.c_size[row] : 0.f;
            inuse[h] = in ? w.c_inuse[row] : 0;
            anc[h] = 0;
            for (int p = parent[h]; p >= 0; p = w.c_parent[o * MAXC + p]) anc[h] |= 1ull << p;
            if (with_body) {
                use_idle[h] = in ? w.c_use_idle[row] : 0.f; use_act[h] = in ? w.c_use_act[row] : 0.f; damage[h] = in ? w.c_damage[row] : 0.f;
                float2 v = in ? reinterpret_cast<const float2 *>(w.c_vel)[row] : make_float2(0.f, 0.f);
                float2 p = in ? reinterpret_cast<const float2 *>(w.c_pos)[row] : make_float2(0.f, 0.f);
                vx[h] = v.x; vy[h] = v.y; px[h] = p.x; py[h] = p.y; av[h] = in ? w.c_angvel[row] : 0.f;
            }
        }
    }
    __device__ uint64_t alive_mask() const {
        uint32_t a0 = __ballot_sync(FULL, alive[0] == SAPIO_CELL_ALIVE), a1 = __ballot_sync(FULL, alive[1] == SAPIO_CELL_ALIVE);
        return ((uint64_t)a1 << 32) | a0;
    }
    __device__ int n_living() const { return __popcll(alive_mask()); }
    __device__ static double dsum(double v) { for (int s = 16; s; s >>= 1) v += __shfl_xor_sync(FULL, v, s); return v; }
    __device__ float sum_living(const float v[2]) const { return warp_sum((alive[0] == 1 ? v[0] : 0.f) + (alive[1] == 1 ? v[1] : 0.f)); }
    __device__ float sum_all(const float v[2]) const { return warp_sum(v[0] + v[1]); }
    // Organism.health_percent / energy_percent (sprites.py:1306-1324)
    __device__ float health_percent() const { float cur = sum_living(health), mx = sum_all(maxhealth); return (cur == 0.f || mx == 0.f) ? 0.f : cur / mx * 100.0f; }
    __device__ float energy_percent() const { return (float)energy_percent_d(); }
    __device__ double max_energy_d() const { return dsum((alive[0] == 1 ? (double)storage[0] : 0.0) + (alive[1] == 1 ? (double)storage[1] : 0.0)); }
    __device__ double total_energy_d() const { return dsum((alive[0] == 1 ? energy[0] : 0.0) + (alive[1] == 1 ? energy[1] : 0.0)); }
    // double versions for the dopamine / pain signal: it is a difference of two nearly equal percentages (sprites.py:1610-1611)
    __device__ double health_percent_d() const {
        double cur = dsum((alive[0] == 1 ? (double)health[0] : 0.0) + (alive[1] == 1 ? (double)health[1] : 0.0)), mx = dsum((double)maxhealth[0] + (double)maxhealth[1]);
        return (cur == 0.0 || mx == 0.0) ? 0.0 : cur / mx * 100.0;
    }
    __device__ double energy_percent_d() const {
        double t = total_energy_d();
        double m = max_energy_d();
        return (t == 0.0 || m == 0.0) ? 0.0 : t / m * 100.0;
    }

    // Organism.kill_cell (sprites.py:1326-1354): the cell and its whole growth sub-tree die; the root of the cascade
    // damages its living parent by break_damage * size. Returns the number of cells that died.
    __device__ int kill(const SapioWorld &w, int k) {
        int pk = bcast2i(parent[0], parent[1], k);
        float sz = bcast2(size[0], size[1], k);
        int n = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            int c = lane + 32 * h;
            bool die = alive[h] == SAPIO_CELL_ALIVE && (c == k || ((anc[h] >> k) & 1ull));
            if (die) { alive[h] = SAPIO_CELL_DYING; health[h] = 0.f; }
            n += __popc(__ballot_sync(FULL, die));
        }
        if (pk >= 0 && lane == (pk & 31)) {
            if (pk >> 5) { if (alive[1] == SAPIO_CELL_ALIVE) health[1] -= w.p.break_damage * sz; }
            else if (alive[0] == SAPIO_CELL_ALIVE) health[0] -= w.p.break_damage * sz;
        }
        return n;
    }
    // Organism.add_energy (sprites.py:1380-1398): equal shares in row order, overflow carried to the next living cell.
    // The carry is a max-plus recurrence c' = max(0, c + e + share - cap): composed with a warp scan of (A, S) pairs,
    // f(c) = max(A, c + S). Returns the overflow of the last living cell.
    // The scan is carried in double: its partial sums of (energy + share - cap) reach 1e4 over 64 cells while the carry they
    // cancel to is O(1) -- in fp32 the refund and the energies behind it were off by 1e-3 (absolute) in well-fed organisms.
    __device__ double add_energy(double amount, int n) {
        if (!n) return 0.0;
        const double per = amount / (double)n;
        double carry = 0.0;
        {   // nobody overflows (the usual case: only a well-fed organism is full): the carry stays zero, no scan
            const bool over = (alive[0] == SAPIO_CELL_ALIVE && energy[0] + per > (double)storage[0]) || (alive[1] == SAPIO_CELL_ALIVE && energy[1] + per > (double)storage[1]);
            if (!__any_sync(FULL, over)) {
#pragma unroll
                for (int h = 0; h < 2; h++) if (alive[h] == SAPIO_CELL_ALIVE) energy[h] += per;
                return 0.0;
            }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            bool a = alive[h] == SAPIO_CELL_ALIVE;
            double A = a ? 0.0 : -INFINITY, S = a ? (energy[h] + per - (double)storage[h]) : 0.0;
            double iA = A, iS = S;                      // inclusive scan over lanes
            for (int d = 1; d < 32; d <<= 1) {
                double pA = __shfl_up_sync(FULL, iA, d), pS = __shfl_up_sync(FULL, iS, d);
                if (lane >= d) { iA = fmax(iA, pA + iS); iS = pS + iS; }
            }
            double eA = __shfl_up_sync(FULL, iA, 1), eS = __shfl_up_sync(FULL, iS, 1);   // exclusive
            if (lane == 0) { eA = -INFINITY; eS = 0.0; }
            double cin = fmax(eA, carry + eS);
            if (a) { double e = energy[h] + per + cin; energy[h] = e > (double)storage[h] ? (double)storage[h] : e; }
            double lA = __shfl_sync(FULL, iA, 31), lS = __shfl_sync(FULL, iS, 31);
            carry = fmax(lA, carry + lS);
        }
        return carry;
    }
    // Organism.take_energy (sprites.py:1400-1407)
    __device__ void take_energy_sum(double per) {
        if (per == 0.0) return;
#pragma unroll
        for (int h = 0; h < 2; h++) if (alive[h] == SAPIO_CELL_ALIVE) energy[h] = fmax(energy[h] - per, 0.0);
    }
    __device__ void store(const SapioWorld &w, bool with_body) const {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            int k = lane + 32 * h, row = o * MAXC + k;
            if (k >= nc) continue;
            w.c_health[row] = health[h]; w.c_energy[row] = (float)energy[h]; w.c_alive[row] = (uint8_t)alive[h];
            if (with_body) {
                w.c_inuse[row] = (uint8_t)inuse[h];
                reinterpret_cast<float2 *>(w.c_vel)[row] = make_float2(vx[h], vy[h]); w.c_angvel[row] = av[h];
            }
        }
    }
};

__device__ __forceinline__ void stat_add(const SapioWorld &w, int which, int n) { if (n) atomicAdd(reinterpret_cast<unsigned long long *>(w.g_stats + which), (unsigned long long)n); }

// Ordered composition of the organisms' gas maps (contract C4) as a scan with a FIXED association: chunk-local scans
// (cub::BlockScan, one CTA per GAS_CHUNK maps), the chunk aggregates composed in order by one thread, the prefix applied.
// cub::DeviceScan's decoupled look-back groups the partial products by arrival time, and the composition of doubles is only
// approximately associative: its results differ in the last bit from run to run, which a world that promises to be
// reproducible from its seed cannot have.
#define GAS_THREADS 256
#define GAS_ITEMS 4
#define GAS_CHUNK (GAS_THREADS * GAS_ITEMS)
__global__ void __launch_bounds__(GAS_THREADS) k_gas_local(const Aff *in, Aff *out, Aff *agg, int n) {
    typedef cub::BlockScan<Aff, GAS_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    const int base = blockIdx.x * GAS_CHUNK + threadIdx.x * GAS_ITEMS;
    Aff v[GAS_ITEMS];
#pragma unroll
    for (int i = 0; i < GAS_ITEMS; i++) v[i] = base + i < n ? in[base + i] : Aff{1.0, 0.0};
    Aff total;
    Scan(tmp).InclusiveScan(v, v, AffCompose(), total);
#pragma unroll
    for (int i = 0; i < GAS_ITEMS; i++) if (base + i < n) out[base + i] = v[i];
    if (threadIdx.x == 0) agg[blockIdx.x] = total;
}
__global__ void k_gas_chunks(Aff *agg, int nchunks) {          // agg[c] becomes the composition of the chunks before c
    if (blockIdx.x || threadIdx.x) return;
    Aff run{1.0, 0.0};
    for (int c = 0; c < nchunks; c++) { const Aff a = agg[c]; agg[c] = run; run = AffCompose()(run, a); }
}
__global__ void __launch_bounds__(GAS_THREADS) k_gas_apply(Aff *out, const Aff *agg, int n) {
    const int c = blockIdx.x + 1;                               // chunk 0 has nothing before it
    const Aff pre = agg[c];
    for (int i = threadIdx.x; i < GAS_CHUNK; i += GAS_THREADS) { const int k = c * GAS_CHUNK + i; if (k < n) out[k] = AffCompose()(pre, out[k]); }
}

#define RULES_WARPS 8
__global__ void __launch_bounds__(RULES_WARPS * 32) k_rules_begin(WORLD_ARG, RulesWs R, uint32_t phases) {
    const SapioParams &p = w.p;
    const int32_t T = (int32_t)w.g_tick[0];
    const float dt = p.dt_wall;
    const double Ttot = w.g_co2[0] + w.g_o2[0];
    const int warps = gridDim.x * RULES_WARPS, lane = lane_id();
    for (int o = blockIdx.x * RULES_WARPS + (threadIdx.x >> 5); o < p.n_org; o += warps) {
        if (w.org_state[o] != 1) { if (lane == 0) { R.gas_map[o] = Aff{1.0, 0.0}; w.org_req[o] = 0; } continue; }
        OrgR g; g.load(w, o, false);
        // exchange rate of a cell (sprites.py:1413-1416), two double divisions: every lane computes its own cells' once, the loop broadcasts
        double rho2[2];
#pragma unroll
        for (int h = 0; h < 2; h++) rho2[h] = (w.org_flags[o] & SAPIO_F_GHOST) ? 0.0 : (double)g.size[h] / ((double)p.width + (double)p.height) * 100.0 * (double)dt / 100.0;   // a ghost's gas exchange belongs to its owner's slab
        const bool immortal = w.org_flags[o] & SAPIO_F_IMMORTAL;
        int died = 0, req = 0;
        if (T - w.org_birth_tick[o] >= p.age_ticks && !immortal) died += g.kill(w, 0);          // sprites.py:1826-1827
        float consumed = w.org_consumed[o];
        if (g.n_living() > 0) {
            float emax = g.sum_living(g.storage);                                                // digestion, sprites.py:1631-1640
            if (consumed >= emax / 2) consumed = emax / 2;
            if (consumed < 0.001f) consumed = 0.f;
            if (consumed != 0.f) {
                float dig = 95.0f / 100.0f * consumed * dt;
                consumed -= dig;
                consumed += (float)g.add_energy((double)dig * 1.2, g.n_living());
            }
            if ((phases & SAPIO_PH_THINK) && w.org_nlayers[o] > 0 && T - w.org_think_tick[o] >= p.think_ticks) {
                double ddt = (double)(T - w.org_dopa_tick[o]) * (double)dt;                       // _update_brain, sprites.py:1608-1625
                double hd = g.health_percent_d() - (double)w.org_last_health[o], ed = g.energy_percent_d() - (double)w.org_last_energy[o];
                double pain = hd / ddt * 5.0;
                if (pain > 0.0) pain = 0.0;
                pain = fabs(pain);
                if (lane == 0) {
                    w.org_energy_diff[o] = (float)(ed / ddt); w.org_pain[o] = (float)pain; w.org_dopamine[o] = (float)(ed / ddt - pain);
                    w.org_dopa_tick[o] = T; w.org_think_tick[o] = T;
                    R.think_list[atomicAdd(&R.counters[RC_NTHINK], 1)] = o;
                }
                req |= SAPIO_REQ_THOUGHT;
            }
        }
        // gas map: simulate which cells reach their turn alive (death checks use start-of-tick health only), then compose
        uint64_t sim = g.alive_mask();
        if (__popcll(sim) == 1) sim = 0;                                                         // only the heart left: it is killed before the loop
        Aff m{1.0, 0.0};
        for (int k = 0; k < g.nc; k++) {
            if (!((sim >> k) & 1ull)) continue;
            float hk = bcast2(g.health[0], g.health[1], k);
            if (hk <= 0.f && !immortal) {
                uint32_t d0 = __ballot_sync(FULL, (g.anc[0] >> k) & 1ull), d1 = __ballot_sync(FULL, (g.anc[1] >> k) & 1ull);
                sim &= ~((((uint64_t)d1 << 32) | d0) | (1ull << k));
                continue;
            }
            double rho = __shfl_sync(FULL, (k >> 5) ? rho2[1] : rho2[0], k & 31);
            int tk = bcast2i(g.type[0], g.type[1], k);
            if (tk == SAPIO_T_CO2C) { m.a = (1.0 - rho) * m.a; m.b = (1.0 - rho) * m.b; }         // co2 -> o2
            else { m.a = (1.0 - rho) * m.a; m.b = (1.0 - rho) * m.b + rho * Ttot; }             // o2 -> co2, o2 = Ttot - co2
        }
        g.store(w, false);
        if (lane == 0) { R.gas_map[o] = m; w.org_consumed[o] = consumed; w.org_req[o] = (uint8_t)req; stat_add(w, SAPIO_STAT_CELL_DEATHS, died); if (req) stat_add(w, SAPIO_STAT_THINKS, 1); }
    }
}

#ifndef RULES_MAIN_MINB
#define RULES_MAIN_MINB 3       // resident CTAs per SM the register allocation aims for: the per-cell loop is a dependent shuffle chain, warps in flight hide it
#endif
__global__ void __launch_bounds__(RULES_WARPS * 32, RULES_MAIN_MINB) k_rules_main(WORLD_ARG, RulesWs R, uint32_t phases) {
    const SapioParams &p = w.p;
    const int32_t T = (int32_t)w.g_tick[0];
    const float dt = p.dt_wall;
    const double co2_0 = w.g_co2[0], Ttot = co2_0 + w.g_o2[0];
    const int NC = p.n_org * MAXC, ND = p.n_dead;
    const int warps = gridDim.x * RULES_WARPS, lane = lane_id();
    int org_ticks = 0;
    for (int o = blockIdx.x * RULES_WARPS + (threadIdx.x >> 5); o < p.n_org; o += warps) {
        if (w.org_state[o] != 1) { if (lane == 0) w.org_gas_refund[o] = 0.0; continue; }
        OrgR g; g.load(w, o, true);
        // exchange rate of a cell (sprites.py:1413-1416), two double divisions: every lane computes its own cells' once, the loop broadcasts
        const bool ghost = w.org_flags[o] & SAPIO_F_GHOST;
        const double rho_0 = ghost ? 0.0 : (double)g.size[0] / ((double)p.width + (double)p.height) * 100.0 * (double)dt / 100.0;
        const double rho_1 = ghost ? 0.0 : (double)g.size[1] / ((double)p.width + (double)p.height) * 100.0 * (double)dt / 100.0;
        const bool immortal = w.org_flags[o] & SAPIO_F_IMMORTAL;
        Aff pre = o > 0 ? R.gas_scan[o - 1] : Aff{1.0, 0.0};
        const double co2_in = p.slab_mode ? R.slab_co2_in[0] : co2_0;                            // slab mode: the slabs before this one have had their turn
        double x = pre.a * co2_in + pre.b, refund = 0.0;                                          // co2 level met by this organism
        const double x_in = x;
        int died = 0, req = w.org_req[o];
        float consumed = w.org_consumed[o];
        float mv0 = w.org_move[4 * o], mv1 = w.org_move[4 * o + 1], mv2 = w.org_move[4 * o + 2], mv3 = w.org_move[4 * o + 3];
        if (g.n_living() > 0) {
            float lh = g.health_percent(), le = g.energy_percent();                              // sprites.py:1653-1654
            if (g.n_living() == 1) died += g.kill(w, 0);                                         // sprites.py:1656-1657
            if (lane == 0 && g.alive[0] == SAPIO_CELL_ALIVE) { w.org_pos[2 * o] = g.px[0]; w.org_pos[2 * o + 1] = g.py[0]; }
            const double emax = g.max_energy_d();
            if (g.n_living() > 0 && g.total_energy_d() >= emax / 1.01 && T - w.org_lastbirth_tick[o] >= p.repro_ticks) {
                uint32_t r[4];
                rng4(p.seed, w.org_uid[o], T, RNG_REPRO_GATE, 0, r);
                float u = u01(r[0]);
                if ((double)u * 100.0 <= g.health_percent_d() && !(w.g_population[0] >= p.population_limit)) {
                    if ((phases & SAPIO_PH_REPRODUCE) && !ghost) req |= SAPIO_REQ_REPRODUCE;
                    if (lane == 0) w.org_lastbirth_tick[o] = T;                                  // sprites.py:1442
                }
#pragma unroll
                for (int h = 0; h < 2; h++) if (g.alive[h] == SAPIO_CELL_ALIVE) g.energy[h] = g.energy[h] / 2;   // sprites.py:1672-1675
                lh = g.health_percent(); le = g.energy_percent();
            }
            if (lane == 0) { w.org_last_health[o] = lh; w.org_last_energy[o] = le; }
            int nliv = g.n_living();                                                             // kept current across kills below
            // take_energy of every cell's usage from every living cell (sprites.py:1400-1407, O(C^2) in the reference): the shares are
            // uniform, and max(max(e - a, 0) - b, 0) == max(e - a - b, 0), so they are summed and taken when somebody looks
            double pending = 0.0;
            uint64_t rot_am = 0; double push_c = 1.0, push_s = 0.0;                              // unit vector of the push direction, cached per set of living cells
            // Organism.rotation (sprites.py:1179-1201) is only used as an angle: cur - old (the % 360 cancels in cos/sin)
            for (int k = 0; k < g.nc; k++) {                                                     // _update_cell, sprites.py:1470-1606
                float hk = bcast2(g.health[0], g.health[1], k);
                int ak = bcast2i(g.alive[0], g.alive[1], k);
                if (hk <= 0.f && !immortal) { if (ak == SAPIO_CELL_ALIVE) { const int nd = g.kill(w, k); died += nd; nliv -= nd; } continue; }
                if (ak != SAPIO_CELL_ALIVE) continue;
                const int tk = bcast2i(g.type[0], g.type[1], k), hsel = k >> 5;
                const bool mine = lane == (k & 31);
                int inuse_k = bcast2i(g.inuse[0], g.inuse[1], k);
                if (tk == SAPIO_T_CARNIV) {                                                      // sprites.py:1491-1516, contracts C2/C3
                    float dmg = bcast2(g.damage[0], g.damage[1], k), add = 0.f;
                    int row = o * MAXC + k, s = w.xa_off[row], e = w.xa_off[row + 1];
                    bool any = false;
                    for (int q = s + lane; q < e; q += 32) {
                        uint32_t other = w.xa_other[q];
                        if ((int)other < NC) { if ((int)(other >> 6) != o && w.c_alive[other] != 0) { any = true; add += dmg / 2; } }
                        else if ((int)other < NC + ND) { int d = (int)other - NC; if (w.d_state[d]) { any = true; add += dmg / 2 * w.d_mass0[d] * w.d_size[d]; } }
                    }
                    consumed += warp_sum(add);
                    inuse_k = __any_sync(FULL, any) ? 1 : 0;
                }
                if (tk == SAPIO_T_PUSH) {                                                        // sprites.py:1519-1544
                    if (fabsf(mv1) >= 0.1f) {
                        // In double like the reference's Python floats, rounded once into the fp32 velocity: a push that is one ulp off
                        // flips the rounding of the integrated position (ulp 0.002 px at x = 2e4, 0.016 px at 2e5), and the pin joints turn
                        // that into 0.01 .. 0.08 px/s of bias velocity -- single-tick parity at 10k organisms and more hangs on this.
                        const uint64_t am = g.alive_mask() & ~1ull;
                        if (am != rot_am) {                                                      // Organism.rotation (sprites.py:1179-1201)
                            // The push direction is atan2(dir_y, dir_x) + rotation, rotation = angle(heart -> first other living cell now) -
                            // angle(the same cell in the genome's frame). Its cosine and sine follow from dot and cross products of the three
                            // vectors -- the angle itself is never needed, and the four double-precision atan2 / sin / cos per organism were a
                            // quarter of this kernel's instructions. Agrees with the trigonometric form to the last bits of a double.
                            rot_am = am;
                            const int rel = __ffsll((long long)am) - 1, rrow = o * MAXC + rel;
                            const double rx = (double)bcast2(g.px[0], g.px[1], rel) - (double)bcast2(g.px[0], g.px[1], 0), ry = (double)bcast2(g.py[0], g.py[1], rel) - (double)bcast2(g.py[0], g.py[1], 0);
                            const double ox = (double)w.c_rel[2 * rrow], oy = (double)w.c_rel[2 * rrow + 1];
                            const double nr = sqrt(rx * rx + ry * ry), no = sqrt(ox * ox + oy * oy), nm = sqrt((double)mv2 * (double)mv2 + (double)mv3 * (double)mv3);
                            const double cx = nr > 0.0 ? rx / nr : 1.0, cy = nr > 0.0 ? ry / nr : 0.0;        // atan2(0, 0) = 0 (a difference of positions is never -0)
                            const double qx = no > 0.0 ? ox / no : 1.0, qy = no > 0.0 ? oy / no : 0.0;
                            // a command rounded to -0.0 keeps its sign (neural.py:292): atan2(+-0, -0) = +-pi, atan2(+-0, +0) = +-0
                            const double mx = nm > 0.0 ? (double)mv2 / nm : (signbit(mv2) ? -1.0 : 1.0), my = nm > 0.0 ? (double)mv3 / nm : 0.0;
                            const double rc = cx * qx + cy * qy, rs = cy * qx - cx * qy;                     // cos, sin of the rotation
                            push_c = mx * rc - my * rs; push_s = my * rc + mx * rs;
                        }
                        if (mine) {
                            const double speed = (double)mv1 * (double)p.max_push * (double)dt;
                            const double dvx = push_c * speed, dvy = push_s * speed;
                            if (hsel) { g.vx[1] = (float)((double)g.vx[1] + dvx); g.vy[1] = (float)((double)g.vy[1] + dvy); }
                            else { g.vx[0] = (float)((double)g.vx[0] + dvx); g.vy[0] = (float)((double)g.vy[0] + dvy); }
                        }
                        inuse_k = 1;
                    } else inuse_k = 0;
                }
                if (tk == SAPIO_T_ROTATE) {                                                      // sprites.py:1546-1552: per tick, not per second
                    if (mine) { if (hsel) g.av[1] = (float)((double)g.av[1] + (double)mv0 * (double)p.max_rot); else g.av[0] = (float)((double)g.av[0] + (double)mv0 * (double)p.max_rot); }
                    inuse_k = fabsf(mv0) >= 0.1f;
                }
                if (mine) { if (hsel) g.inuse[1] = inuse_k; else g.inuse[0] = inuse_k; }
                float idle = bcast2(g.use_idle[0], g.use_idle[1], k), act = bcast2(g.use_act[0], g.use_act[1], k); double usage;   // sprites.py:1554-1574
                if (inuse_k) {
                    if (tk == SAPIO_T_PUSH || tk == SAPIO_T_ROTATE) { double e = ((double)act - (double)idle) * (double)(tk == SAPIO_T_PUSH ? mv1 : mv0); usage = e < 0.0 ? fabs(-(double)idle + e) : (double)idle + e; }
                    else usage = (double)act;
                } else usage = (double)idle;
                pending += usage * (double)dt / (double)nliv;
                const double rho = __shfl_sync(FULL, hsel ? rho_1 : rho_0, k & 31);
                if (tk == SAPIO_T_CO2C) {                                                        // convert_co2, sprites.py:1409-1428 (C4)
                    double gq = rho * x;
                    x -= gq;
                    g.take_energy_sum(pending); pending = 0.0;
                    const double spare = g.add_energy(gq * (double)p.co2_ratio, nliv);
                    refund += spare / (double)p.co2_ratio;
                } else { double gq = rho * (Ttot - x); x += gq; }                               // convert_o2, sprites.py:1430-1439
                if (mine) {
                    // static indices only: a runtime index would move these per-lane arrays from registers to local memory
                    float cap = hsel ? g.storage[1] : g.storage[0], mh = hsel ? g.maxhealth[1] : g.maxhealth[0], hh = hsel ? g.health[1] : g.health[0];
                    double e = fmax((hsel ? g.energy[1] : g.energy[0]) - pending, 0.0);
                    if (e > (double)cap) { e = (double)cap; if (hsel) g.energy[1] = e + pending; else g.energy[0] = e + pending; }
                    if (e >= (double)cap / 2 && hh < mh) { hh += p.heal_rate / 100.0f * mh * dt; if (hh > mh) hh = mh; }        // sprites.py:1589-1594
                    if (e <= 2.0 / 100.0 * (double)cap) hh -= p.starvation_rate / 100.0f * mh * dt;                           // sprites.py:1603-1606
                    if (hsel) g.health[1] = hh; else g.health[0] = hh;
                }
            }
            g.take_energy_sum(pending);
            if ((phases & SAPIO_PH_TRAIN) && !(w.org_flags[o] & SAPIO_F_MALADAPTIVE) && w.org_nlayers[o] > 0 && T - w.org_train_tick[o] >= p.train_ticks) req |= SAPIO_REQ_TRAIN;
            org_ticks += !ghost;                                                                 // counted per warp, added once (124k atomics on one address otherwise)
        }
        g.store(w, true);
        if (lane == 0) {
            w.org_consumed[o] = consumed; w.org_req[o] = (uint8_t)req; w.org_gas_in[o] = x_in; w.org_gas_refund[o] = refund;
            stat_add(w, SAPIO_STAT_CELL_DEATHS, died);
            if (req & SAPIO_REQ_TRAIN) R.train_list[atomicAdd(&R.counters[RC_NTRAIN], 1)] = o;
        }
    }
    if (lane == 0) stat_add(w, SAPIO_STAT_ORG_TICKS, org_ticks);
}

// gases after the rules phase: the composed map of every organism, plus the refunds handed back at the end (C4)
__global__ void k_rules_finish(WORLD_ARG, RulesWs R) {
    if (w.p.slab_mode) return;                               // the host composes the slabs' maps and refunds (slabs.py)
    const double co2_0 = w.g_co2[0], Ttot = co2_0 + w.g_o2[0];
    Aff tot = R.gas_scan[w.p.n_org - 1];
    double co2 = tot.a * co2_0 + tot.b + R.refund_sum[0];
    w.g_co2[0] = co2; w.g_o2[0] = Ttot - co2;
}

// ---- settle ------------------------------------------------------------------------------------------------------
// contract C2: victims gather. A living, non-barrier cell loses the damage of every living carnivore cell of another
// organism it touches; a dead cell loses damage/size of mass per carnivore (sprites.py:1497-1516), attackers in id order.
__global__ void __launch_bounds__(256) k_settle_gather(WORLD_ARG) {
    const int NC = w.p.n_org * MAXC, ND = w.p.n_dead, NB = NC + ND;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < NB; v += gridDim.x * blockDim.x) {
        int s = w.xa_off[v], e = w.xa_off[v + 1];
        if (s == e) continue;
        bool cell = v < NC;
        if (cell) { if (w.c_alive[v] != SAPIO_CELL_ALIVE) continue; } else if (!w.d_state[v - NC]) continue;
        float sum = 0.f; bool hit = false;
        for (int q = s; q < e; q++) {
            uint32_t a = w.xa_other[q];
            if ((int)a >= NC || w.c_alive[a] != SAPIO_CELL_ALIVE || w.c_type[a] != SAPIO_T_CARNIV) continue;
            if (cell && (int)(a >> 6) == (v >> 6)) continue;
            sum += cell ? w.c_damage[a] : w.c_damage[a] / w.c_size[a];
            hit = true;
        }
        if (!hit) continue;
        if (cell) { if (w.c_type[v] != SAPIO_T_BARRIER) w.c_health[v] -= sum; }
        else { float m = w.d_mass[v - NC] - sum; w.d_mass[v - NC] = m > 0.f ? m : 0.f; }
    }
}
// how many cells of each organism died this tick / which dead slots are free
__global__ void __launch_bounds__(256) k_settle_count(WORLD_ARG, RulesWs R) {
    const int n_org = w.p.n_org, ND = w.p.n_dead;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    for (int o = gtid; o <= n_org; o += gthreads) {
        int n = 0;
        if (o < n_org && w.org_state[o] == 1) for (int k = 0; k < w.org_ncells[o]; k++) n += (w.c_alive[o * MAXC + k] == SAPIO_CELL_DYING);
        R.dying_cnt[o] = n;
    }
    for (int d = gtid; d <= ND; d += gthreads) R.free_flag[d] = (d < ND && !w.d_state[d]) ? 1 : 0;
}
__global__ void __launch_bounds__(256) k_settle_freelist(WORLD_ARG, RulesWs R) {
    const int ND = w.p.n_dead;
    for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < ND; d += gridDim.x * blockDim.x) if (R.free_flag[d]) R.free_list[R.free_off[d]] = d;
}
// _make_dead_cell (sprites.py:1164-1177, 1341-1349): the j-th dying cell in (organism, row) order takes the j-th smallest
// free slot; organisms without a living cell leave organism_list (sprites.py:1837)
__global__ void __launch_bounds__(256) k_settle_deaths(WORLD_ARG, RulesWs R) {
    const int n_org = w.p.n_org, ND = w.p.n_dead;
    const int n_free = R.free_off[ND];
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_org; o += gridDim.x * blockDim.x) {
        if (w.org_state[o] != 1) continue;
        int rank = R.dying_off[o], living = 0;
        for (int k = 0; k < w.org_ncells[o]; k++) {
            int row = o * MAXC + k;
            int a = w.c_alive[row];
            living += (a == SAPIO_CELL_ALIVE);
            if (a != SAPIO_CELL_DYING) continue;
            if (rank < n_free) {
                int d = R.free_list[rank];
                w.d_state[d] = 2;                                                                 // created this tick
                reinterpret_cast<float2 *>(w.d_pos)[d] = reinterpret_cast<const float2 *>(w.c_pos)[row];
                reinterpret_cast<float2 *>(w.d_vel)[d] = reinterpret_cast<const float2 *>(w.c_vel)[row];    // sprites.py:1348-1349
                w.d_ang[d] = 0.f; w.d_angvel[d] = w.c_angvel[row];
                reinterpret_cast<float2 *>(w.d_vbias)[d] = make_float2(0.f, 0.f); w.d_wbias[d] = 0.f;
                w.d_mass[d] = w.d_mass0[d] = w.c_mass[row];
                w.d_size[d] = w.c_size[row]; w.d_elast[d] = w.c_elast[row]; w.d_fric[d] = w.c_fric[row];
                w.d_owner[d] = w.org_uid[o];
            } else atomicAdd(reinterpret_cast<unsigned long long *>(w.g_stats + SAPIO_STAT_OVERFLOW), 1ull);
            rank++;
            w.c_alive[row] = 0; w.c_health[row] = 0.f;
        }
        if (!living) { w.org_state[o] = 0; atomicAdd(reinterpret_cast<unsigned long long *>(w.g_stats + SAPIO_STAT_DEATHS), 1ull); }
    }
}
// gradual_dispersion (sprites.py:1752-1772) + bounds (sprites.py:1849-1856). The o2 -> co2 hand-back of the dead cells is
// (co2 - T) *= prod(1 - rate_k/100): the factors commute, so a plain product reduction reproduces the sequential loop.
__global__ void __launch_bounds__(256) k_settle_disperse(WORLD_ARG, RulesWs R) {
    const SapioParams &p = w.p;
    const int NC = p.n_org * MAXC, ND = p.n_dead;
    const float dt = p.dt_wall;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    for (int d = gtid; d < ND; d += gthreads) {
        double factor = 1.0;
        if (w.d_state[d]) {
            float mass = w.d_mass[d];
            bool remove = false;
            if (mass < p.mass_cutoff) remove = true;
            else {
                float rate = p.dispersion_rate / 100.0f * mass * dt, nm = mass - rate;
                if (nm > 0.f) { w.d_mass[d] = nm; factor = (w.p.slab_mode && R.dead_ghost[d]) ? 1.0 : 1.0 - (double)rate / 100.0; } else remove = true;
            }
            float2 q = reinterpret_cast<const float2 *>(w.d_pos)[d];
            if (q.x < 0.f || q.x > p.width || q.y < 0.f || q.y > p.height) remove = true;
            if (remove) { w.d_state[d] = 0; atomicAdd(reinterpret_cast<unsigned long long *>(w.g_stats + SAPIO_STAT_DEAD_REMOVED), 1ull); }
        }
        R.disp_factor[d] = factor;
    }
    for (int row = gtid; row < NC; row += gthreads) {                     // living cells out of bounds: health = 0
        if (w.c_alive[row] != SAPIO_CELL_ALIVE || w.org_state[row >> 6] != 1) continue;
        float2 q = reinterpret_cast<const float2 *>(w.c_pos)[row];
        if (q.x < 0.f || q.x > p.width || q.y < 0.f || q.y > p.height) w.c_health[row] = 0.f;
    }
}
struct MulD { __host__ __device__ double operator()(double a, double b) const { return a * b; } };
__global__ void k_settle_gas(WORLD_ARG, RulesWs R) {
    if (w.p.slab_mode) return;
    const double co2 = w.g_co2[0], Ttot = co2 + w.g_o2[0];
    double c2 = Ttot + (co2 - Ttot) * R.disp_prod[0];
    w.g_co2[0] = c2; w.g_o2[0] = Ttot - c2;
}
