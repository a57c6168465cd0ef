/* sapio_migrate.h -- wire format of a migrating organism (population islands).
 *
 * A record is everything the world holds for one organism slot -- every array of scope ORG / CELL / PAIR / ICON of
 * sapio_fields.def: genome tables, cell state, pin and contact impulses, brain, optimiser state, memories -- cut down to the
 * extents in use: cell rows up to org_ncells, the brain's P parameters, Adam moments of the trained layers, memory rows with
 * their I columns. ~40 kB for the average organism instead of the 431 kB of a full-capacity slot.
 *
 *   buffer  = SapioMigBufHeader | record 0 | record 1 | ...          (records start at hdr.offset[r], 16-byte aligned)
 *   record  = SapioMigHeader | field 0 | field 1 | ...               (fields in sapio_fields.def order, each padded to 16 bytes)
 *
 * Shared by the CUDA library (csrc/migrate.cuh) and the CPU oracle (oracle/migrate.c); plain C.
 * The reference's analogue of a migrating organism is a copied / pickled DNA (Sapiogenesis.py:325-335, 413-424).
 */
#ifndef SAPIO_MIGRATE_H
#define SAPIO_MIGRATE_H
#include <string.h>
#include "sapio.h"
#ifndef MAXC
#define MAXC SAPIO_MAXC               /* the field list writes per-row extents with this name */
#endif

#define SAPIO_MIG_MAX_FIELDS 112
#define SAPIO_MIG_MAX_RECORDS 4096
#define SAPIO_MIG_MAGIC 0x53415031           /* "SAP1" */
#define SAPIO_MIG_CANDIDATES 360             /* placement candidates per arrival, like the 360 bearings of get_new_organism_position */

enum { SAPIO_MIGK_FULL = 0, SAPIO_MIGK_CELLROWS, SAPIO_MIGK_PAIRS, SAPIO_MIGK_ICON, SAPIO_MIGK_BRAIN, SAPIO_MIGK_ADAM,
       SAPIO_MIGK_HIST_IN, SAPIO_MIGK_STM_IN, SAPIO_MIGK_MEM_IN, SAPIO_MIGK_MEM_OUT, SAPIO_MIGK_MEM_REW };

typedef struct { char *base; uint32_t full_bytes; uint32_t unit; int32_t kind; int32_t pad; } SapioMigField;   /* unit: bytes per row / element */
typedef struct {
    SapioMigField f[SAPIO_MIG_MAX_FIELDS];
    int32_t n, in_cap;
    int32_t i_uid, i_cpos, i_cspos, i_orgpos, i_calive, i_csize, i_ncells, i_state;      /* field indices the arrival code needs */
    int32_t i_tick[5];                                                                    /* tick stamps, re-based to the destination's clock */
} SapioMigTable;

typedef struct {            /* 64 bytes, first thing of a record */
    int32_t magic, ncells, nlayers, I, P, npar, mem_n, stm_n, hist_n, ic_n, bytes /* whole record */, src_slot;
    int32_t pad[4];
} SapioMigHeader;

typedef struct {            /* first thing of a buffer */
    int32_t magic, n_records, used_bytes, chosen /* emigrants picked */, stayed /* picked but no room in the buffer */, src_tick, pad[2];
    int32_t offset[SAPIO_MIG_MAX_RECORDS];
} SapioMigBufHeader;

static inline int sapio_mig_kind(const char *name, const char *scope) {
    if (!strcmp(scope, "CELL")) return SAPIO_MIGK_CELLROWS;
    if (!strcmp(scope, "PAIR")) return SAPIO_MIGK_PAIRS;
    if (!strcmp(scope, "ICON")) return SAPIO_MIGK_ICON;
    if (!strcmp(name, "org_brain")) return SAPIO_MIGK_BRAIN;
    if (!strcmp(name, "org_adam_m") || !strcmp(name, "org_adam_v")) return SAPIO_MIGK_ADAM;
    if (!strcmp(name, "org_hist_in")) return SAPIO_MIGK_HIST_IN;
    if (!strcmp(name, "org_stm_in")) return SAPIO_MIGK_STM_IN;
    if (!strcmp(name, "org_mem_in")) return SAPIO_MIGK_MEM_IN;
    if (!strcmp(name, "org_mem_out")) return SAPIO_MIGK_MEM_OUT;
    if (!strcmp(name, "org_mem_rew")) return SAPIO_MIGK_MEM_REW;
    return SAPIO_MIGK_FULL;
}

/* the field table of a world (w may be NULL: sizes only) */
static inline void sapio_mig_table(const SapioParams *p, const SapioWorld *w, SapioMigTable *T) {
    const size_t IN = (size_t)p->in_cap, BRAIN = (size_t)p->brain_cap, ADAM = (size_t)p->adam_cap, RNNH = (size_t)p->rnn_cap;
    (void)IN; (void)BRAIN; (void)ADAM; (void)RNNH;
    const size_t M_ORG = 1, M_CELL = SAPIO_MAXC, M_PAIR = SAPIO_NPAIR, M_ICON = SAPIO_ICAP, M_DEAD = 0, M_XCON = 0, M_BODY1 = 0, M_GLOB = 0;
    T->n = 0; T->in_cap = p->in_cap;
    T->i_uid = T->i_cpos = T->i_cspos = T->i_orgpos = T->i_calive = T->i_csize = T->i_ncells = T->i_state = -1;
    for (int k = 0; k < 5; k++) T->i_tick[k] = -1;
#define SAPIO_FIELD(name, ctype, scope, per) \
    if (M_##scope) { SapioMigField *e = &T->f[T->n]; e->base = w ? (char *)w->name : (char *)0; e->unit = (uint32_t)(sizeof(ctype) * (size_t)(per)); \
                     e->full_bytes = (uint32_t)(e->unit * M_##scope); e->kind = sapio_mig_kind(#name, #scope); e->pad = 0; \
                     if (!strcmp(#name, "org_uid")) T->i_uid = T->n; if (!strcmp(#name, "c_pos")) T->i_cpos = T->n; if (!strcmp(#name, "c_spos")) T->i_cspos = T->n; \
                     if (!strcmp(#name, "org_pos")) T->i_orgpos = T->n; if (!strcmp(#name, "c_alive")) T->i_calive = T->n; if (!strcmp(#name, "c_size")) T->i_csize = T->n; \
                     if (!strcmp(#name, "org_ncells")) T->i_ncells = T->n; if (!strcmp(#name, "org_state")) T->i_state = T->n; \
                     if (!strcmp(#name, "org_birth_tick")) T->i_tick[0] = T->n; if (!strcmp(#name, "org_lastbirth_tick")) T->i_tick[1] = T->n; \
                     if (!strcmp(#name, "org_think_tick")) T->i_tick[2] = T->n; if (!strcmp(#name, "org_train_tick")) T->i_tick[3] = T->n; \
                     if (!strcmp(#name, "org_dopa_tick")) T->i_tick[4] = T->n; T->n++; }
#include "sapio_fields.def"
#undef SAPIO_FIELD
}

#ifdef __CUDACC__
#define SAPIO_HD __host__ __device__
#else
#define SAPIO_HD
#endif

/* bytes of field k that travel for an organism described by h (before padding to 16) */
SAPIO_HD static inline uint32_t sapio_mig_used(const SapioMigField *f, const SapioMigHeader *h) {
    switch (f->kind) {
        case SAPIO_MIGK_CELLROWS: return f->unit * (uint32_t)h->ncells;
        case SAPIO_MIGK_PAIRS: {      /* lexicographic pairs of rows i < j < ncells: a prefix of the slot's table */
            int n = h->ncells; if (n < 2) return 0;
            int last = (n - 2) * (2 * SAPIO_MAXC - (n - 2) - 1) / 2 + ((n - 1) - (n - 2) - 1);
            return f->unit * (uint32_t)(last + 1);
        }
        case SAPIO_MIGK_ICON: return f->unit * (uint32_t)h->ic_n;
        case SAPIO_MIGK_BRAIN: return 4u * (uint32_t)h->P;
        case SAPIO_MIGK_ADAM: return 4u * (uint32_t)h->npar;
        case SAPIO_MIGK_HIST_IN: return 4u * (uint32_t)(h->hist_n * h->I);
        case SAPIO_MIGK_STM_IN: return 4u * (uint32_t)(h->stm_n * h->I);
        case SAPIO_MIGK_MEM_IN: return 4u * (uint32_t)(h->mem_n * h->I);
        case SAPIO_MIGK_MEM_OUT: return 16u * (uint32_t)h->mem_n;
        case SAPIO_MIGK_MEM_REW: return 4u * (uint32_t)h->mem_n;
        default: return f->full_bytes;
    }
}
/* rows of a strided (in_cap-wide) table that travels compacted to I columns; 0: the field is copied flat */
SAPIO_HD static inline int sapio_mig_rows(const SapioMigField *f, const SapioMigHeader *h) {
    switch (f->kind) {
        case SAPIO_MIGK_HIST_IN: return h->hist_n;
        case SAPIO_MIGK_STM_IN: return h->stm_n;
        case SAPIO_MIGK_MEM_IN: return h->mem_n;
        default: return 0;
    }
}
/* the record header of the organism in slot o */
SAPIO_HD static inline void sapio_mig_header(const SapioWorld *w, int o, SapioMigHeader *h) {
    const int32_t *ld = w->org_ldim + 16 * o;
    int nl = w->org_nlayers[o];
    h->magic = SAPIO_MIG_MAGIC; h->ncells = w->org_ncells[o]; h->nlayers = nl; h->src_slot = o;
    h->I = nl > 0 ? ld[0] : 0; h->P = 0; h->npar = 0;
    for (int l = 0; l < nl; l++) h->P += ld[l + 1] * ld[l] + ld[l + 1];
    if (nl > 0) h->npar = ld[1] * ld[0] + ld[1] + 4 * ld[nl - 1] + 4;            /* the input and output layers (networks.py:221) */
    if (nl > 0 && w->org_rnn_h[o] > 0) {                                        /* RNN brain: wider input layer, second head (networks.py:151-166) */
        const int H = w->org_rnn_h[o];
        h->P += H * ld[1] + H * ld[nl - 1] + H;
    }
    h->mem_n = w->org_mem_n[o] < SAPIO_MEMROWS ? w->org_mem_n[o] : SAPIO_MEMROWS;
    h->stm_n = w->org_stm_n[o] < SAPIO_STM ? w->org_stm_n[o] : SAPIO_STM;       /* rings: slots 0 .. min(count, capacity) - 1 are in use */
    h->hist_n = w->org_hist_n[o] < SAPIO_HIST ? w->org_hist_n[o] : SAPIO_HIST;
    h->ic_n = w->org_ic_n[o];
    if (h->mem_n < 0) h->mem_n = 0;
    if (h->stm_n < 0) h->stm_n = 0;
    if (h->hist_n < 0) h->hist_n = 0;
    h->bytes = 0; h->pad[0] = h->pad[1] = h->pad[2] = h->pad[3] = 0;
}
/* bytes of the whole record (header + padded fields) */
SAPIO_HD static inline uint32_t sapio_mig_record_bytes(const SapioMigTable *T, const SapioMigHeader *h) {
    uint32_t b = (uint32_t)sizeof(SapioMigHeader);
    for (int k = 0; k < T->n; k++) b += (sapio_mig_used(&T->f[k], h) + 15u) & ~15u;
    return b;
}
#endif
