// halo.cuh -- spatial slabs of one large world (second mode; BASELINE configs[4]): the halo exchange.
//
// A slab [slab_x0, slab_x1) owns the organisms whose heart cell lies in it and the dead cells whose centre lies in it, and carries
// copies ("ghosts") of what its neighbours own within `halo` of the common boundary, so that everything an owned organism can touch,
// see or smell this tick is resident with the state it had at the start of the tick. Every tick, before the rules:
//   sapio_halo_pack    the owned organisms (whole ragged records, include/sapio_migrate.h) and owned dead cells in the strip next to
//                      a boundary -> buffers for that neighbour (NCCL send/recv by the host, sapiogenesis_b200/slabs.py)
//   sapio_halo_unpack  the neighbour's records -> local slots: the slot a ghost had last tick (direct map by sender slot: contact
//                      caches are keyed by body id), else a free one. uids, positions and clocks are kept -- it is the same organism.
//                      A record whose heart lies in THIS slab is adopted (ownership hand-over), the sender demotes its copy.
//   sapio_halo_sweep   ghosts that were not refreshed left the halo: freed. Owned organisms / dead cells that left the slab become ghosts.
// Ghosts are simulated redundantly for one tick (they collide, are seen, think) and overwritten by their owner's state at the next
// exchange; they take no part in the gas exchange, do not reproduce and are not counted (rules.cuh, SAPIO_F_GHOST).
#pragma once
#include "migrate2.cuh"

struct HaloWs {
    int *ghost_slot[2];      // [n_org]  local slot + 1 of the ghost of the side's neighbour's slot (0: none)
    int *ghost_dslot[2];     // [n_dead] the same for dead cells
    int *refreshed;          // [n_org]  tick + 1 at which the slot's ghost was last refreshed
    int *drefreshed;         // [n_dead]
    int *dcount;             // [8] 0: dead chosen, 1: cursor scratch
    int *dlist;              // [n_dead] chosen dead slots (pack), assigned local slots (unpack)
};
struct DeadRec { float pos[2], vel[2], ang, angvel, vbias[2], wbias, mass, mass0, size, elast, fric; int64_t owner; int32_t state, src_slot; };   // 72 bytes
struct DeadBufHeader { int32_t magic, n, dropped, pad; };

// owned living organisms with the heart in [x_lo, x_hi), ascending slot (deterministic), at most `cap`: flags, an exclusive scan
// (grid.cuh), scatter -- three grid-wide launches instead of one CTA walking every slot
__global__ void __launch_bounds__(256) k_halo_flag(WORLD_ARG, float x_lo, float x_hi, int *__restrict__ flag) {
    const int n_org = w.p.n_org;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o <= n_org; o += gridDim.x * blockDim.x) {
        bool in = false;
        if (o < n_org && w.org_state[o] == 1 && !(w.org_flags[o] & SAPIO_F_GHOST)) { const float hx = w.c_pos[2 * (o * MAXC)]; in = hx >= x_lo && hx < x_hi; }
        flag[o] = in;
    }
}
__global__ void __launch_bounds__(256) k_halo_choose(WORLD_ARG, MigWs M, const int *__restrict__ flag, const int *__restrict__ off, int cap) {
    const int n_org = w.p.n_org;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_org; o += gridDim.x * blockDim.x)
        if (flag[o] && off[o] < cap) M.emig[off[o]] = o;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int total = off[n_org];
        M.counters[MC_CHOSEN] = min(total, cap);
        if (total > cap) atomicAdd((unsigned long long *)&w.g_stats[SAPIO_STAT_OVERFLOW], 1ull);
    }
}
// owned dead cells with the centre in [x_lo, x_hi): records straight into the buffer (order: atomic, the receiver maps by sender slot)
__global__ void __launch_bounds__(256) k_halo_pack_dead(WORLD_ARG, RulesWs R, float x_lo, float x_hi, char *buf, size_t cap_bytes) {
    DeadBufHeader *H = reinterpret_cast<DeadBufHeader *>(buf);
    DeadRec *rec = reinterpret_cast<DeadRec *>(buf + sizeof(DeadBufHeader));
    const int cap = (int)((cap_bytes - sizeof(DeadBufHeader)) / sizeof(DeadRec));
    for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < w.p.n_dead; d += gridDim.x * blockDim.x) {
        if (!w.d_state[d] || R.dead_ghost[d]) continue;
        const float x = w.d_pos[2 * d];
        if (!(x >= x_lo && x < x_hi)) continue;
        const int k = atomicAdd(&H->n, 1);
        if (k >= cap) { atomicAdd(&H->dropped, 1); continue; }
        DeadRec r;
        r.pos[0] = x; r.pos[1] = w.d_pos[2 * d + 1]; r.vel[0] = w.d_vel[2 * d]; r.vel[1] = w.d_vel[2 * d + 1]; r.ang = w.d_ang[d]; r.angvel = w.d_angvel[d];
        r.vbias[0] = w.d_vbias[2 * d]; r.vbias[1] = w.d_vbias[2 * d + 1]; r.wbias = w.d_wbias[d]; r.mass = w.d_mass[d]; r.mass0 = w.d_mass0[d];
        r.size = w.d_size[d]; r.elast = w.d_elast[d]; r.fric = w.d_fric[d]; r.owner = w.d_owner[d]; r.state = w.d_state[d]; r.src_slot = d;
        rec[k] = r;
    }
}
__global__ void k_halo_dead_header(char *buf) { DeadBufHeader *H = reinterpret_cast<DeadBufHeader *>(buf); H->magic = SAPIO_MIG_MAGIC; H->n = 0; H->dropped = 0; H->pad = 0; }

// ---- arriving: slots. A ghost that is still here keeps last tick's slot (same uid): a thread per record. The others -- new ghosts,
// organisms that changed their mapping -- are served by one warp in record order: the slot that already holds the uid, else the highest
// free slot (deterministic; a handful per tick, where the serial form walked ~900 records with four dependent global loads each).
__global__ void __launch_bounds__(256) k_halo_assign_kept(WORLD_ARG, const __grid_constant__ SapioMigTable T, MigWs M, HaloWs Hw, int side, const char *buf) {
    const SapioMigBufHeader *H = reinterpret_cast<const SapioMigBufHeader *>(buf);
    const int n = H->magic == SAPIO_MIG_MAGIC ? min(H->n_records, SAPIO_MIG_MAX_RECORDS) : 0;
    const int stamp = (int)w.g_tick[0] + 1;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const char *rec = buf + H->offset[r];
        const SapioMigHeader h = *reinterpret_cast<const SapioMigHeader *>(rec);
        const int64_t uid = *reinterpret_cast<const int64_t *>(rec + mig_field_offset(T, h, T.i_uid));
        int slot = Hw.ghost_slot[side][h.src_slot] - 1;
        if (slot >= 0 && !(w.org_state[slot] == 1 && w.org_uid[slot] == uid)) slot = -1;
        M.slot[r] = slot >= 0 ? slot : -2;                           // -2: k_halo_assign decides
        if (slot >= 0) Hw.refreshed[slot] = stamp;
    }
}
__global__ void __launch_bounds__(32) k_halo_assign(WORLD_ARG, const __grid_constant__ SapioMigTable T, MigWs M, HaloWs Hw, int side, const char *buf) {
    const SapioMigBufHeader *H = reinterpret_cast<const SapioMigBufHeader *>(buf);
    const int n = H->magic == SAPIO_MIG_MAGIC ? min(H->n_records, SAPIO_MIG_MAX_RECORDS) : 0;
    const int lane = lane_id();
    int cursor = w.p.n_org - 1, n_acc = 0, n_drop = 0;
    for (int r0 = 0; r0 < n; r0 += 32) {
        const int mine = r0 + lane < n ? M.slot[r0 + lane] : 0;
        n_acc += __popc(__ballot_sync(FULL, r0 + lane < n && mine >= 0));
        uint32_t pending = __ballot_sync(FULL, r0 + lane < n && mine == -2);
        while (pending) {
            const int r = r0 + __ffs(pending) - 1; pending &= pending - 1;
            const char *rec = buf + H->offset[r];
            const SapioMigHeader h = *reinterpret_cast<const SapioMigHeader *>(rec);
            const int64_t uid = *reinterpret_cast<const int64_t *>(rec + mig_field_offset(T, h, T.i_uid));
            int slot = -1;
            // the same organism may already live here under another mapping (it was ours, or came from the other side): look for its uid
            for (int base = 0; base < w.p.n_org && slot < 0; base += 32) {
                const int o = base + lane;
                const uint32_t hit = __ballot_sync(FULL, o < w.p.n_org && w.org_state[o] == 1 && w.org_uid[o] == uid);
                if (hit) slot = base + __ffs(hit) - 1;
            }
            if (slot < 0) {
                while (cursor >= 0) {                               // highest free slot: the owned organisms keep the low ones
                    const int o = cursor - lane;
                    const uint32_t fb = __ballot_sync(FULL, o >= 0 && w.org_state[o] == 0 && Hw.refreshed[o] != (int)w.g_tick[0] + 1);
                    if (fb) { slot = cursor - (__ffs(fb) - 1); cursor = slot - 1; break; }
                    cursor -= 32;
                }
            }
            if (lane == 0) {
                M.slot[r] = slot;
                if (slot >= 0) { Hw.ghost_slot[side][h.src_slot] = slot + 1; Hw.refreshed[slot] = (int)w.g_tick[0] + 1; }
            }
            if (slot >= 0) n_acc++; else n_drop++;
            __syncwarp();
        }
    }
    if (lane == 0) { M.counters[MC_ACCEPTED] = n_acc; M.counters[MC_DROPPED] = n_drop; if (n_drop) atomicAdd((unsigned long long *)&w.g_stats[SAPIO_STAT_OVERFLOW], (unsigned long long)n_drop); }
}
// records into their slots, as they are (same organism, same clock); ghost unless its heart lies in this slab
__global__ void __launch_bounds__(256) k_halo_unpack(WORLD_ARG, const __grid_constant__ SapioMigTable T, MigWs M, char *buf) {
    const SapioMigBufHeader *H = reinterpret_cast<const SapioMigBufHeader *>(buf);
    const int n = H->magic == SAPIO_MIG_MAGIC ? min(H->n_records, SAPIO_MIG_MAX_RECORDS) : 0;
    __shared__ SapioMigHeader h;
    for (int r = blockIdx.x; r < n; r += gridDim.x) {
        const int slot = M.slot[r];
        if (slot < 0) continue;
        char *rec = buf + H->offset[r];
        __syncthreads();
        if (threadIdx.x == 0) h = *reinterpret_cast<SapioMigHeader *>(rec);
        __syncthreads();
        uint32_t off = sizeof(SapioMigHeader);
        for (int k = 0; k < T.n; k++) {
            mig_field<false>(T.f[k], h, T.in_cap, T.f[k].base + (size_t)slot * T.f[k].full_bytes, rec + off);
            off += (sapio_mig_used(&T.f[k], &h) + 15u) & ~15u;
        }
        __syncthreads();
        uint8_t *al = reinterpret_cast<uint8_t *>(T.f[T.i_calive].base + (size_t)slot * T.f[T.i_calive].full_bytes);
        for (int k = h.ncells + threadIdx.x; k < MAXC; k += blockDim.x) al[k] = 0;
        if (threadIdx.x == 0) {
            const float hx = w.c_pos[2 * (slot * MAXC)];
            const bool mine = hx >= w.p.slab_x0 && hx < w.p.slab_x1;
            w.org_flags[slot] = (uint8_t)((w.org_flags[slot] & ~SAPIO_F_GHOST) | (mine ? 0 : SAPIO_F_GHOST));
            w.org_req[slot] = 0;
        }
    }
}
__global__ void __launch_bounds__(32) k_halo_assign_dead(WORLD_ARG, RulesWs R, HaloWs Hw, int side, const char *buf) {
    const DeadBufHeader *H = reinterpret_cast<const DeadBufHeader *>(buf);
    const DeadRec *rec = reinterpret_cast<const DeadRec *>(buf + sizeof(DeadBufHeader));
    const int n = H->magic == SAPIO_MIG_MAGIC ? H->n - H->dropped : 0;
    const int lane = lane_id(), stamp = (int)w.g_tick[0] + 1;
    int cursor = w.p.n_dead - 1;
    for (int r0 = 0; r0 < n; r0 += 32) {                              // mapped records: a lane each; the others allocate one after the other
        const int r = r0 + lane;
        int slot = -1;
        if (r < n) { slot = Hw.ghost_dslot[side][rec[r].src_slot] - 1; if (slot >= 0 && !(w.d_state[slot] && R.dead_ghost[slot])) slot = -1; }
        uint32_t need = __ballot_sync(FULL, r < n && slot < 0);
        while (need) {
            const int src = __ffs(need) - 1; need &= need - 1;
            int got = -1;
            while (cursor >= 0) {
                const int d = cursor - lane;
                const uint32_t fb = __ballot_sync(FULL, d >= 0 && w.d_state[d] == 0 && Hw.drefreshed[d] != stamp);
                if (fb) { got = cursor - (__ffs(fb) - 1); cursor = got - 1; break; }
                cursor -= 32;
            }
            if (lane == src) slot = got;
        }
        if (r < n) {
            Hw.dlist[r] = slot;
            if (slot >= 0) { Hw.ghost_dslot[side][rec[r].src_slot] = slot + 1; Hw.drefreshed[slot] = stamp; }
            else atomicAdd((unsigned long long *)&w.g_stats[SAPIO_STAT_OVERFLOW], 1ull);
        }
        __syncwarp();
    }
}
__global__ void __launch_bounds__(256) k_halo_unpack_dead(WORLD_ARG, RulesWs R, HaloWs Hw, const char *buf) {
    const DeadBufHeader *H = reinterpret_cast<const DeadBufHeader *>(buf);
    const DeadRec *rec = reinterpret_cast<const DeadRec *>(buf + sizeof(DeadBufHeader));
    const int n = H->magic == SAPIO_MIG_MAGIC ? H->n - H->dropped : 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const int d = Hw.dlist[r];
        if (d < 0) continue;
        const DeadRec q = rec[r];
        w.d_pos[2 * d] = q.pos[0]; w.d_pos[2 * d + 1] = q.pos[1]; w.d_vel[2 * d] = q.vel[0]; w.d_vel[2 * d + 1] = q.vel[1]; w.d_ang[d] = q.ang; w.d_angvel[d] = q.angvel;
        w.d_vbias[2 * d] = q.vbias[0]; w.d_vbias[2 * d + 1] = q.vbias[1]; w.d_wbias[d] = q.wbias; w.d_mass[d] = q.mass; w.d_mass0[d] = q.mass0;
        w.d_size[d] = q.size; w.d_elast[d] = q.elast; w.d_fric[d] = q.fric; w.d_owner[d] = q.owner; w.d_state[d] = (uint8_t)q.state; w.d_eaten[d] = 0.f;
        R.dead_ghost[d] = !(q.pos[0] >= w.p.slab_x0 && q.pos[0] < w.p.slab_x1);
    }
}
// after both sides: what was not refreshed left the halo; what left the slab is a ghost from now on
__global__ void __launch_bounds__(256) k_halo_sweep(WORLD_ARG, RulesWs R, HaloWs Hw) {
    const int stamp = (int)w.g_tick[0] + 1;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    for (int o = gtid; o < w.p.n_org; o += gthreads) {
        if (w.org_state[o] != 1) continue;
        if (w.org_flags[o] & SAPIO_F_GHOST) {
            if (Hw.refreshed[o] != stamp) {                          // released like an emigrant (k_mig_release), rows emptied
                w.org_state[o] = 0; w.org_req[o] = 0; w.org_ic_n[o] = 0; w.org_ncells[o] = 0;
                for (int k = 0; k < MAXC; k++) w.c_alive[o * MAXC + k] = 0;
            }
        } else {
            const float hx = w.c_pos[2 * (o * MAXC)];
            if (!(hx >= w.p.slab_x0 && hx < w.p.slab_x1)) { w.org_flags[o] |= SAPIO_F_GHOST; Hw.refreshed[o] = stamp; }   // its new owner refreshes it from the next tick on
        }
    }
    for (int d = gtid; d < w.p.n_dead; d += gthreads) {
        if (!w.d_state[d]) { R.dead_ghost[d] = 0; continue; }
        if (R.dead_ghost[d]) { if (Hw.drefreshed[d] != stamp) { w.d_state[d] = 0; R.dead_ghost[d] = 0; } }
        else { const float x = w.d_pos[2 * d]; if (!(x >= w.p.slab_x0 && x < w.p.slab_x1)) { R.dead_ghost[d] = 1; Hw.drefreshed[d] = stamp; } }
    }
}
// host-composed gas scalars (slabs.py allgathers the slabs' maps; co2 is the same number on every slab)
__global__ void k_slab_gas_begin(WORLD_ARG, RulesWs R, const double *maps /* [n][2] */, int rank) {
    double x = w.g_co2[0];
    for (int k = 0; k < rank; k++) x = maps[2 * k] * x + maps[2 * k + 1];
    R.slab_co2_in[0] = x;
}
__global__ void k_slab_gas_finish(WORLD_ARG, const double *maps, int n, const double *refunds /* [n] */) {
    const double co2_0 = w.g_co2[0], Ttot = co2_0 + w.g_o2[0];
    double x = co2_0, ref = 0.0;
    for (int k = 0; k < n; k++) { x = maps[2 * k] * x + maps[2 * k + 1]; ref += refunds[k]; }
    x += ref;
    w.g_co2[0] = x; w.g_o2[0] = Ttot - x;
}
__global__ void k_slab_disperse_finish(WORLD_ARG, const double *prods, int n) {
    const double co2 = w.g_co2[0], Ttot = co2 + w.g_o2[0];
    double pr = 1.0;
    for (int k = 0; k < n; k++) pr *= prods[k];
    const double c2 = Ttot + (co2 - Ttot) * pr;
    w.g_co2[0] = c2; w.g_o2[0] = Ttot - c2;
}
__global__ void k_slab_population(WORLD_ARG, const int *counts, int n) { int s = 0; for (int k = 0; k < n; k++) s += counts[k]; w.g_population[0] = s; }
