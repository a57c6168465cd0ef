// migrate.cuh -- whole-organism records for island migration (population islands, one per GPU).
// An emigrant is everything the world holds per organism slot -- every array of scope ORG / CELL / PAIR / ICON in
// include/sapio_fields.def: genome tables, cell state, pin and contact impulses, brain, optimizer state, memories -- packed
// into one contiguous record. Records travel between ranks (NCCL send/recv, sapiogenesis_b200/islands.py) and are
// scattered into free slots of the destination world, which gives them fresh uids (the Philox key of their later draws).
// The field table is generated from the X-macro list, so a new field migrates without touching this file.
#pragma once
#include "common.cuh"

#define MIG_MAX_FIELDS 112
struct MigField { char *base; uint32_t bytes; uint32_t off; };          // bytes per organism, offset inside a record
struct MigTable { MigField f[MIG_MAX_FIELDS]; int n; uint32_t record_bytes; uint32_t uid_off; };

static inline void mig_table(const SapioParams &p, const SapioWorld *w, MigTable &T) {
    const size_t IN = (size_t)p.in_cap, BRAIN = (size_t)p.brain_cap, ADAM = (size_t)p.adam_cap, RNNH = (size_t)p.rnn_cap;
    (void)IN; (void)BRAIN; (void)ADAM; (void)RNNH;
    const size_t M_ORG = 1, M_CELL = MAXC, M_PAIR = SAPIO_NPAIR, M_ICON = SAPIO_ICAP, M_DEAD = 0, M_XCON = 0, M_BODY1 = 0, M_GLOB = 0;
    T.n = 0; uint32_t off = 0; T.uid_off = 0;
#define SAPIO_FIELD(name, ctype, scope, per) \
    if (M_##scope) { MigField &e = T.f[T.n++]; e.base = w ? (char *)w->name : nullptr; e.bytes = (uint32_t)(sizeof(ctype) * (size_t)(per) * M_##scope); \
                     e.off = off; if (!strcmp(#name, "org_uid")) T.uid_off = off; off += (e.bytes + 15u) & ~15u; }
#include "../../include/sapio_fields.def"
#undef SAPIO_FIELD
    T.record_bytes = off;
}

__device__ __forceinline__ void mig_copy(char *dst, const char *src, uint32_t bytes) {
    if ((((uintptr_t)dst | (uintptr_t)src | bytes) & 15u) == 0) {
        for (uint32_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
    } else if ((((uintptr_t)dst | (uintptr_t)src | bytes) & 3u) == 0) {
        for (uint32_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(dst)[i] = reinterpret_cast<const uint32_t *>(src)[i];
    } else {
        for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
    }
}

// one CTA per record. PACK: world -> record. !PACK: record -> world slot, uid = uid0 + record index.
template <bool PACK>
__global__ void __launch_bounds__(256) k_org_records(const __grid_constant__ MigTable T, const int32_t *__restrict__ slots, int n, char *buf,
                                                      const int64_t *next_uid) {
    for (int r = blockIdx.x; r < n; r += gridDim.x) {
        const int slot = slots[r];
        char *rec = buf + (size_t)r * T.record_bytes;
        for (int k = 0; k < T.n; k++) {
            char *field = T.f[k].base + (size_t)slot * T.f[k].bytes;
            if (PACK) mig_copy(rec + T.f[k].off, field, T.f[k].bytes);
            else if (T.f[k].off == T.uid_off) { if (threadIdx.x == 0) *reinterpret_cast<int64_t *>(field) = next_uid[0] + r; }
            else mig_copy(field, rec + T.f[k].off, T.f[k].bytes);
        }
    }
}
__global__ void k_uid_advance(int64_t *next_uid, int n) { next_uid[0] += n; }
// emigrants leave without a trace: slot free, rows empty (no dead bodies: the organism lives on elsewhere)
__global__ void __launch_bounds__(64) k_release_slots(WORLD_ARG, const int32_t *__restrict__ slots, int n) {
    for (int r = blockIdx.x; r < n; r += gridDim.x) {
        const int slot = slots[r];
        w.c_alive[slot * MAXC + threadIdx.x] = 0;
        if (threadIdx.x == 0) { w.org_state[slot] = 0; w.org_req[slot] = 0; w.org_ic_n[slot] = 0; w.org_ncells[slot] = 0; }
    }
}
// cached cross-contact impulses are keyed by body id (slot * 64 + row): a contact of a departed organism must not warm-start
// whoever moves into its slot
__global__ void __launch_bounds__(256) k_release_contacts(WORLD_ARG) {
    const int NC = w.p.n_org * MAXC, nx = min(w.g_x_n[0], w.p.x_cap);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nx; c += gridDim.x * blockDim.x) {
        const uint64_t key = w.x_key[c];
        const uint32_t a = (uint32_t)(key >> 32), b = (uint32_t)key;
        if ((a < (uint32_t)NC && w.org_state[a >> 6] == 0) || (b < (uint32_t)NC && w.org_state[b >> 6] == 0)) { w.x_jn[c] = 0.f; w.x_jt[c] = 0.f; }
    }
}
