// migrate2.cuh -- island migration on the device, end to end (first mode): choice of the emigrants, ragged
// records (include/sapio_migrate.h), release of their slots, and on the other side placement of the arrivals by the reference's
// rule for a new organism -- a candidate position is viable when the organism's box lies inside the world and blocks no living
// organism's box (viable_organism_position, sprites.py:156-192), dead cells inside the accepted box are removed
// (sprites.py:184-190) -- then the scatter into free slots. No host round trip anywhere: the buffer has a fixed size, counts
// travel in its header.
//
//   sapio_migrate_out : k_mig_choose -> k_mig_sizes -> k_mig_pack -> k_mig_release -> k_mig_scrub
//   sapio_migrate_in  : k_org_rects (residents) -> k_mig_prepare -> k_mig_viable -> k_mig_resolve -> k_mig_unpack -> k_mig_clear_dead
//
// Order is part of the contract (oracle/migrate.c restates it): emigrants are the first `count` living slots met on the cyclic
// walk start, start + 7919, ... (start from Philox(seed; -2 - rank, tick, RNG_MIGRATE)); arrivals are placed in record order, the
// j-th accepted one takes the j-th smallest free slot and uid next_uid + j; candidate k of record r is the k-th Philox draw of
// (seed; next_uid + r, tick, RNG_MIGRATE), uniform over the world.
#pragma once
#include "common.cuh"
#include "rng.cuh"
#include "../../include/sapio_migrate.h"

#define MIG_STRIDE 7919
struct MigWs {
    int *emig;            // [MAX_RECORDS] chosen slots
    int *counters;        // 0 chosen, 1 packed, 2 accepted, 3 dropped (no viable place / no free slot), 4..7 totals since the world began
    double *rect;         // [MAX_RECORDS][6] box relative to the heart (4), heart position (2)
    uint32_t *mask;       // [MAX_RECORDS][12] viable candidates among the first upto[r] (the rest: not evaluated)
    int *upto;            // [MAX_RECORDS] candidates evaluated by k_mig_viable (a multiple of 32, or all of them)
    int *slot;            // [MAX_RECORDS] destination slot or -1
    double *delta;        // [MAX_RECORDS][2] translation
    double *acc_rect;     // [MAX_RECORDS][4] boxes of the accepted arrivals (absolute)
};
enum { MC_CHOSEN = 0, MC_PACKED = 1, MC_ACCEPTED = 2, MC_DROPPED = 3, MC_TOT_OUT = 4, MC_TOT_IN = 5, MC_TOT_DROPPED = 6, MC_TOT_STAYED = 7 };

// ---- leaving ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mig_choose(WORLD_ARG, MigWs M, int count, int rank) {
    __shared__ int s_base, s_warp[8];
    const int n_org = w.p.n_org, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t r[4];
    rng4(w.p.seed, (int64_t)(-2 - rank), (int32_t)w.g_tick[0], RNG_MIGRATE, 0, r);
    const int start = (int)(r[0] % (uint32_t)n_org);
    const int stride = (n_org % MIG_STRIDE) ? MIG_STRIDE : 7907;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n_org; i0 += 256) {
        const int i = i0 + tid;
        int slot = -1;
        if (i < n_org) { slot = (int)(((long long)start + (long long)i * stride) % n_org); if (w.org_state[slot] != 1) slot = -1; }
        const uint32_t b = __ballot_sync(FULL, slot >= 0);
        if (lane == 0) s_warp[wid] = __popc(b);
        __syncthreads();
        int before = s_base;
        for (int k = 0; k < wid; k++) before += s_warp[k];
        const int pos = before + __popc(b & ((1u << lane) - 1u));
        if (slot >= 0 && pos < count) M.emig[pos] = slot;
        __syncthreads();
        if (tid == 0) { int t = s_base; for (int k = 0; k < 8; k++) t += s_warp[k]; s_base = t; }
        __syncthreads();
        if (s_base >= count) break;
    }
    if (tid == 0) M.counters[MC_CHOSEN] = min(s_base, count);
}
// record sizes, offsets in order, the cut at the buffer's capacity
__global__ void __launch_bounds__(256) k_mig_sizes(WORLD_ARG, const __grid_constant__ SapioMigTable T, MigWs M, char *buf, size_t cap) {
    __shared__ int s_bytes[SAPIO_MIG_MAX_RECORDS];
    const int n = M.counters[MC_CHOSEN];
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        SapioMigHeader h; sapio_mig_header(&w, M.emig[r], &h);
        s_bytes[r] = (int)sapio_mig_record_bytes(&T, &h);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        SapioMigBufHeader *H = reinterpret_cast<SapioMigBufHeader *>(buf);
        size_t off = (sizeof(SapioMigBufHeader) + 15) & ~size_t(15);
        int k = 0;
        for (; k < n; k++) { if (off + (size_t)s_bytes[k] > cap) break; H->offset[k] = (int32_t)off; off += (size_t)s_bytes[k]; }
        H->magic = SAPIO_MIG_MAGIC; H->n_records = k; H->used_bytes = (int32_t)off; H->chosen = n; H->stayed = n - k; H->src_tick = (int32_t)w.g_tick[0];
        H->pad[0] = H->pad[1] = 0;
        M.counters[MC_PACKED] = k; M.counters[MC_TOT_OUT] += k; M.counters[MC_TOT_STAYED] += n - k;
    }
}
__device__ __forceinline__ void mig_copy_flat(char *dst, const char *src, uint32_t bytes) {
    if ((((uintptr_t)dst | (uintptr_t)src | bytes) & 15u) == 0) {
        for (uint32_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
    } else if ((((uintptr_t)dst | (uintptr_t)src | bytes) & 3u) == 0) {
        for (uint32_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(dst)[i] = reinterpret_cast<const uint32_t *>(src)[i];
    } else {
        for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
    }
}
// one field between a world slot and a record; strided tables travel compacted to their I columns
template <bool PACK>
__device__ __forceinline__ void mig_field(const SapioMigField &f, const SapioMigHeader &h, int in_cap, char *field /* world, this slot */, char *rec /* in the record */) {
    const uint32_t used = sapio_mig_used(&f, &h);
    const int rows = sapio_mig_rows(&f, &h);
    if (rows == 0) { if (PACK) mig_copy_flat(rec, field, used); else mig_copy_flat(field, rec, used); return; }
    const int I = h.I;
    float *wf = reinterpret_cast<float *>(field), *rf = reinterpret_cast<float *>(rec);
    for (int e = threadIdx.x; e < rows * I; e += blockDim.x) {
        const int rr = e / I, c = e - rr * I;
        if (PACK) rf[e] = wf[(size_t)rr * in_cap + c]; else wf[(size_t)rr * in_cap + c] = rf[e];
    }
}
__global__ void __launch_bounds__(256) k_mig_pack(WORLD_ARG, const __grid_constant__ SapioMigTable T, MigWs M, char *buf) {
    const SapioMigBufHeader *H = reinterpret_cast<const SapioMigBufHeader *>(buf);
    __shared__ SapioMigHeader h;
    for (int r = blockIdx.x; r < H->n_records; r += gridDim.x) {
        const int slot = M.emig[r];
        char *rec = buf + H->offset[r];
        __syncthreads();
        if (threadIdx.x == 0) { sapio_mig_header(&w, slot, &h); h.bytes = (int32_t)sapio_mig_record_bytes(&T, &h); *reinterpret_cast<SapioMigHeader *>(rec) = h; }
        __syncthreads();
        uint32_t off = sizeof(SapioMigHeader);
        for (int k = 0; k < T.n; k++) {
            mig_field<true>(T.f[k], h, T.in_cap, T.f[k].base + (size_t)slot * T.f[k].full_bytes, rec + off);
            off += (sapio_mig_used(&T.f[k], &h) + 15u) & ~15u;
        }
    }
}
// emigrants leave without a trace: slot free, rows empty (no dead bodies: the organism lives on elsewhere), their own adjacency emptied
__global__ void __launch_bounds__(64) k_mig_release(WORLD_ARG, const int *__restrict__ slots, const int *n_ptr, int n_direct) {
    const int n = n_ptr ? *n_ptr : n_direct;
    const int na = w.xa_off[w.p.n_org * MAXC + w.p.n_dead];
    for (int r = blockIdx.x; r < n; r += gridDim.x) {
        const int slot = slots[r], row = slot * MAXC + threadIdx.x;
        w.c_alive[row] = 0;
        for (int q = w.xa_off[row], e = min(w.xa_off[row + 1], na); q < e; q++) w.xa_other[q] = 0x7FFFFFFFu;   // nobody (above every body id, also as a signed int)
        if (threadIdx.x == 0) { w.org_state[slot] = 0; w.org_req[slot] = 0; w.org_ic_n[slot] = 0; w.org_ncells[slot] = 0; }
    }
}
// Cached cross-contact impulses and the adjacency are keyed by body id (slot * 64 + row): what pointed at a departed organism must
// neither warm-start nor be eaten / bitten by whoever moves into its slot before the next Space.step rebuilds both lists.
__global__ void __launch_bounds__(256) k_mig_scrub(WORLD_ARG) {
    const int NC = w.p.n_org * MAXC, nx = min(w.g_x_n[0], w.p.x_cap);
    const int na = min(w.xa_off[NC + w.p.n_dead], 2 * w.p.x_cap);
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    for (int c = gtid; c < nx; c += gthreads) {
        const uint64_t key = w.x_key[c];
        const uint32_t a = (uint32_t)(key >> 32), b = (uint32_t)key;
        if ((a < (uint32_t)NC && w.org_state[a >> 6] == 0) || (b < (uint32_t)NC && w.org_state[b >> 6] == 0)) { w.x_jn[c] = 0.f; w.x_jt[c] = 0.f; }
    }
    for (int q = gtid; q < na; q += gthreads) {
        const uint32_t o = w.xa_other[q];
        if (o < (uint32_t)NC && w.org_state[o >> 6] == 0) w.xa_other[q] = 0x7FFFFFFFu;
    }
}

// ---- arriving -----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mig_field_offset(const SapioMigTable &T, const SapioMigHeader &h, int idx) {
    uint32_t off = sizeof(SapioMigHeader);
    for (int k = 0; k < idx; k++) off += (sapio_mig_used(&T.f[k], &h) + 15u) & ~15u;
    return off;
}
// the arriving organism's box relative to its heart (Organism.get_rect, sprites.py:1211-1232, on the cells of the record)
__global__ void __launch_bounds__(128) k_mig_prepare(WORLD_ARG, const __grid_constant__ SapioMigTable T, MigWs M, const char *buf) {
    const SapioMigBufHeader *H = reinterpret_cast<const SapioMigBufHeader *>(buf);
    const int n = H->magic == SAPIO_MIG_MAGIC ? min(H->n_records, SAPIO_MIG_MAX_RECORDS) : 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) { M.counters[MC_ACCEPTED] = 0; M.counters[MC_DROPPED] = 0; }
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const char *rec = buf + H->offset[r];
        const SapioMigHeader h = *reinterpret_cast<const SapioMigHeader *>(rec);
        const float *pos = reinterpret_cast<const float *>(rec + mig_field_offset(T, h, T.i_cpos));
        const float *size = reinterpret_cast<const float *>(rec + mig_field_offset(T, h, T.i_csize));
        const uint8_t *alive = reinterpret_cast<const uint8_t *>(rec + mig_field_offset(T, h, T.i_calive));
        double minx = pos[0], maxx = minx, miny = pos[1], maxy = miny;
        for (int k = 0; k < h.ncells; k++) {
            if (alive[k] != SAPIO_CELL_ALIVE) continue;
            double x = pos[2 * k], y = pos[2 * k + 1], rad = (double)size[k] / 2.0;
            if (x - rad < minx) minx = x - rad;
            if (x + rad > maxx) maxx = x + rad;
            if (y - rad < miny) miny = y - rad;
            if (y + rad > maxy) maxy = y + rad;
        }
        const double hx = pos[0], hy = pos[1];
        double *R = M.rect + 6 * r;
        R[0] = minx - hx; R[1] = miny - hy; R[2] = maxx - hx; R[3] = maxy - hy; R[4] = hx; R[5] = hy;
    }
}
__device__ __forceinline__ bool mig_rects_block(const double d[4], const double r[4]) {       // == rects_block of repro.cuh (sprites.py:180-182)
    bool inx = (d[0] > r[0] && d[0] < r[2]) || (d[2] > r[0] && d[2] < r[2]) || (r[0] > d[0] && r[2] < d[2]);
    if (!inx) return false;
    return (d[1] > r[1] && d[1] < r[3]) || (d[3] > r[1] && d[3] < r[3]) || (r[1] > d[1] && r[3] < d[3]);
}
__device__ __forceinline__ void mig_candidate(const SapioWorld &w, int r, int k, double *x, double *y) {
    uint32_t q[4];
    rng4(w.p.seed, w.g_next_uid[0] + r, (int32_t)w.g_tick[0], RNG_MIGRATE, (uint32_t)k, q);
    *x = (double)u01(q[0]) * (double)w.p.width; *y = (double)u01(q[1]) * (double)w.p.height;
}
// One candidate of arrival r against the world's bounds and every living organism's box, by the lanes of a warp (strided over the
// organisms). A conservative fp32 test (boxes rounded outwards) rejects almost every pair; the exact double predicate decides the rest.
__device__ __forceinline__ bool mig_blocked_by_residents(const SapioWorld &w, const double d[4], int first, int stride) {
    const float f0 = __double2float_rd(d[0]), f1 = __double2float_rd(d[1]), f2 = __double2float_ru(d[2]), f3 = __double2float_ru(d[3]);
    bool blocked = false;
    for (int o = first; o < w.p.n_org && !blocked; o += stride) {
        if (w.org_state[o] != 1) continue;
        const double *R = w.org_rect + 4 * o;
        const double r0 = R[0], r1 = R[1], r2 = R[2], r3 = R[3];
        if (f2 < __double2float_rd(r0) || f0 > __double2float_ru(r2) || f3 < __double2float_rd(r1) || f1 > __double2float_ru(r3)) continue;
        const double rr[4] = {r0, r1, r2, r3};
        blocked = mig_rects_block(d, rr);
    }
    return blocked;
}
// Candidates are tried in order and the first viable one wins (get_new_organism_position, sprites.py:194-219), so they are evaluated
// lazily: 32 at a time (MIG_VTHREADS / 32 slices of the organisms per candidate), stopping after the first round that holds a viable
// one. In a world that is 5 % covered that is the first round; k_mig_resolve evaluates further candidates itself in the rare case
// that every evaluated one is blocked by an earlier arrival. (All 360 x every organism used to cost 24 ms per exchange.)
#define MIG_VTHREADS 512
__global__ void __launch_bounds__(MIG_VTHREADS) k_mig_viable(WORLD_ARG, MigWs M, const char *buf) {
    const SapioMigBufHeader *H = reinterpret_cast<const SapioMigBufHeader *>(buf);
    const int n = H->magic == SAPIO_MIG_MAGIC ? min(H->n_records, SAPIO_MIG_MAX_RECORDS) : 0;
    __shared__ int blocked_by[32];
    __shared__ uint32_t round_mask;
    const int cand = threadIdx.x & 31, slice = threadIdx.x >> 5, slices = MIG_VTHREADS / 32;
    for (int r = blockIdx.x; r < n; r += gridDim.x) {
        const double *R = M.rect + 6 * r;
        int upto = 0;
        for (int k0 = 0; k0 < SAPIO_MIG_CANDIDATES; k0 += 32) {
            __syncthreads();
            if (threadIdx.x < 32) blocked_by[threadIdx.x] = 0;
            __syncthreads();
            const int k = k0 + cand;
            bool ok = k < SAPIO_MIG_CANDIDATES;
            double d[4] = {0, 0, 0, 0};
            if (ok) {
                double x, y; mig_candidate(w, r, k, &x, &y);
                d[0] = R[0] + x; d[1] = R[1] + y; d[2] = R[2] + x; d[3] = R[3] + y;
                ok = !(d[0] < 0 || d[2] > (double)w.p.width || d[1] < 0 || d[3] > (double)w.p.height);         // sprites.py:170-171
            }
            if (ok && mig_blocked_by_residents(w, d, slice, slices)) blocked_by[cand] = 1;
            __syncthreads();
            if (threadIdx.x < 32) {
                const uint32_t m = __ballot_sync(FULL, ok && !blocked_by[cand]);
                if (threadIdx.x == 0) { M.mask[12 * r + (k0 >> 5)] = m; round_mask = m; }
            }
            __syncthreads();
            upto = min(k0 + 32, SAPIO_MIG_CANDIDATES);
            if (round_mask) break;
        }
        if (threadIdx.x == 0) M.upto[r] = upto;
    }
}
// in record order: the first viable candidate that no earlier arrival's box blocks; the j-th accepted takes the j-th smallest free slot
__global__ void __launch_bounds__(32) k_mig_resolve(WORLD_ARG, MigWs M, const char *buf) {
    const SapioMigBufHeader *H = reinterpret_cast<const SapioMigBufHeader *>(buf);
    const int n = H->magic == SAPIO_MIG_MAGIC ? min(H->n_records, SAPIO_MIG_MAX_RECORDS) : 0;
    const int lane = lane_id();
    int n_acc = 0, n_drop = 0, cursor = 0;
    for (int r = 0; r < n; r++) {
        const double *R = M.rect + 6 * r;
        const int upto = M.upto[r];
        int chosen = -1; double cx = 0, cy = 0;
        for (int k = 0; k < SAPIO_MIG_CANDIDATES && chosen < 0; k++) {
            double x, y; mig_candidate(w, r, k, &x, &y);
            const double d[4] = {R[0] + x, R[1] + y, R[2] + x, R[3] + y};
            if (k < upto) { if (!((M.mask[12 * r + (k >> 5)] >> (k & 31)) & 1u)) continue; }
            else {                                                   // beyond what k_mig_viable evaluated: the warp does it (rare)
                if (d[0] < 0 || d[2] > (double)w.p.width || d[1] < 0 || d[3] > (double)w.p.height) continue;
                if (__any_sync(FULL, mig_blocked_by_residents(w, d, lane, 32))) continue;
            }
            bool blocked = false;
            for (int a = lane; a < n_acc; a += 32) blocked |= mig_rects_block(d, M.acc_rect + 4 * a);
            if (!__any_sync(FULL, blocked)) { chosen = k; cx = x; cy = y; }
        }
        int slot = -1;
        if (chosen >= 0) {                                           // next free slot at or after the cursor
            while (cursor < w.p.n_org) {
                const int o = cursor + lane;
                const uint32_t fb = __ballot_sync(FULL, o < w.p.n_org && w.org_state[o] == 0);
                if (fb) { slot = cursor + __ffs(fb) - 1; cursor = slot + 1; break; }
                cursor += 32;
            }
        }
        if (lane == 0) {
            M.slot[r] = slot;
            if (slot >= 0) {
                M.delta[2 * r] = cx - R[4]; M.delta[2 * r + 1] = cy - R[5];
                double *A = M.acc_rect + 4 * n_acc;
                A[0] = R[0] + cx; A[1] = R[1] + cy; A[2] = R[2] + cx; A[3] = R[3] + cy;
            }
        }
        if (slot >= 0) n_acc++; else n_drop++;
        __syncwarp();
    }
    if (lane == 0) { M.counters[MC_ACCEPTED] = n_acc; M.counters[MC_DROPPED] = n_drop; M.counters[MC_TOT_IN] += n_acc; M.counters[MC_TOT_DROPPED] += n_drop; }
}
__global__ void __launch_bounds__(256) k_mig_unpack(WORLD_ARG, const __grid_constant__ SapioMigTable T, MigWs M, char *buf) {
    const SapioMigBufHeader *H = reinterpret_cast<const SapioMigBufHeader *>(buf);
    const int n = H->magic == SAPIO_MIG_MAGIC ? min(H->n_records, SAPIO_MIG_MAX_RECORDS) : 0;
    __shared__ SapioMigHeader h;
    __shared__ int s_rank;
    const int32_t dtick = (int32_t)w.g_tick[0] - H->src_tick;
    for (int r = blockIdx.x; r < n; r += gridDim.x) {
        const int slot = M.slot[r];
        if (slot < 0) continue;
        char *rec = buf + H->offset[r];
        __syncthreads();
        if (threadIdx.x == 0) { h = *reinterpret_cast<SapioMigHeader *>(rec); int k = 0; for (int q = 0; q < r; q++) k += M.slot[q] >= 0; s_rank = k; }
        __syncthreads();
        uint32_t off = sizeof(SapioMigHeader);
        for (int k = 0; k < T.n; k++) {
            mig_field<false>(T.f[k], h, T.in_cap, T.f[k].base + (size_t)slot * T.f[k].full_bytes, rec + off);
            off += (sapio_mig_used(&T.f[k], &h) + 15u) & ~15u;
        }
        __syncthreads();
        // what the new world decides: uid, position, clock; rows beyond the organism's are empty
        const double dx = M.delta[2 * r], dy = M.delta[2 * r + 1];
        float *cp = reinterpret_cast<float *>(T.f[T.i_cpos].base + (size_t)slot * T.f[T.i_cpos].full_bytes);
        float *sp = reinterpret_cast<float *>(T.f[T.i_cspos].base + (size_t)slot * T.f[T.i_cspos].full_bytes);
        uint8_t *al = reinterpret_cast<uint8_t *>(T.f[T.i_calive].base + (size_t)slot * T.f[T.i_calive].full_bytes);
        for (int k = threadIdx.x; k < MAXC; k += blockDim.x) {
            if (k >= h.ncells) { al[k] = 0; continue; }
            cp[2 * k] = (float)((double)cp[2 * k] + dx); cp[2 * k + 1] = (float)((double)cp[2 * k + 1] + dy);
            sp[2 * k] = (float)((double)sp[2 * k] + dx); sp[2 * k + 1] = (float)((double)sp[2 * k + 1] + dy);
        }
        if (threadIdx.x == 0) {
            float *op = reinterpret_cast<float *>(T.f[T.i_orgpos].base + (size_t)slot * T.f[T.i_orgpos].full_bytes);
            op[0] = (float)((double)op[0] + dx); op[1] = (float)((double)op[1] + dy);
            *reinterpret_cast<int64_t *>(T.f[T.i_uid].base + (size_t)slot * T.f[T.i_uid].full_bytes) = w.g_next_uid[0] + s_rank;
            for (int q = 0; q < 5; q++) *reinterpret_cast<int32_t *>(T.f[T.i_tick[q]].base + (size_t)slot * T.f[T.i_tick[q]].full_bytes) += dtick;
            w.org_req[slot] = 0;
        }
    }
}
// dead cells whose centre lies strictly inside an accepted box are deleted (sprites.py:184-190, 215-217)
__global__ void __launch_bounds__(256) k_mig_clear_dead(WORLD_ARG, MigWs M) {
    const int n_acc = M.counters[MC_ACCEPTED];
    if (n_acc == 0) return;
    int removed = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w.p.n_dead; i += gridDim.x * blockDim.x) {
        if (!w.d_state[i]) continue;
        const double x = w.d_pos[2 * i], y = w.d_pos[2 * i + 1];
        for (int a = 0; a < n_acc; a++) {
            const double *d = M.acc_rect + 4 * a;
            if (x > d[0] && x < d[2] && y > d[1] && y < d[3]) { w.d_state[i] = 0; removed++; break; }
        }
    }
    if (removed) atomicAdd(reinterpret_cast<unsigned long long *>(w.g_stats + SAPIO_STAT_DEAD_REMOVED), (unsigned long long)removed);
}
__global__ void k_mig_finish(WORLD_ARG, MigWs M) { w.g_next_uid[0] += M.counters[MC_ACCEPTED]; }
