/* migrate.c -- island migration, CPU restatement of sapiogenesis_b200/csrc/migrate2.cuh (TEST INFRASTRUCTURE, see oracle.h).
 *
 * The reference has one world; its analogue of a migrant is a copied / pickled DNA (Sapiogenesis.py:325-335, 413-424) dropped at
 * the mouse position. What this restates from the reference is the placement rule: a new organism goes where its box lies inside
 * the world and blocks no living organism's box (viable_organism_position, sprites.py:156-192), and dead cells inside the box are
 * deleted (sprites.py:184-190, 215-217). Everything else here is OUR contract (order of choices, wire format:
 * include/sapio_migrate.h), implemented a second time in plain loops so that the CUDA path can be compared with it bit for bit. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"
#include "../include/sapio_migrate.h"

static void copy_field(int pack, const SapioMigField *f, const SapioMigHeader *h, int in_cap, char *field, char *rec) {
    uint32_t used = sapio_mig_used(f, h);
    int rows = sapio_mig_rows(f, h);
    if (rows == 0) { if (pack) memcpy(rec, field, used); else memcpy(field, rec, used); return; }
    float *wf = (float *)field, *rf = (float *)rec;
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < h->I; c++) { if (pack) rf[r * h->I + c] = wf[(size_t)r * in_cap + c]; else wf[(size_t)r * in_cap + c] = rf[r * h->I + c]; }
}

static void scrub(SapioWorld *w) {
    const int NC = w->p.n_org * MAXC, nx = w->g_x_n[0] < w->p.x_cap ? w->g_x_n[0] : w->p.x_cap;
    int na = w->xa_off[NC + w->p.n_dead]; if (na > 2 * w->p.x_cap) na = 2 * w->p.x_cap;
    for (int c = 0; c < nx; c++) {
        uint32_t a = (uint32_t)(w->x_key[c] >> 32), b = (uint32_t)w->x_key[c];
        if ((a < (uint32_t)NC && w->org_state[a >> 6] == 0) || (b < (uint32_t)NC && w->org_state[b >> 6] == 0)) { w->x_jn[c] = 0.f; w->x_jt[c] = 0.f; }
    }
    for (int q = 0; q < na; q++) { uint32_t o = w->xa_other[q]; if (o < (uint32_t)NC && w->org_state[o >> 6] == 0) w->xa_other[q] = 0x7FFFFFFFu; }
}

/* picks, packs, releases: returns the number of records written, or < 0 */
int oracle_migrate_out(SapioWorld *w, int count, int rank, void *buf_, size_t cap) {
    char *buf = (char *)buf_;
    const int n_org = w->p.n_org;
    if (count > SAPIO_MIG_MAX_RECORDS || cap < sizeof(SapioMigBufHeader) + 64) return SAPIO_E_ARG;
    SapioMigTable T; sapio_mig_table(&w->p, w, &T);
    uint32_t r[4];
    sapio_rng4(w->p.seed, (int64_t)(-2 - rank), (int32_t)w->g_tick[0], SAPIO_RNG_MIGRATE, 0, r);
    const int start = (int)(r[0] % (uint32_t)n_org), stride = (n_org % 7919) ? 7919 : 7907;
    int *emig = (int *)malloc(sizeof(int) * (size_t)(count > 0 ? count : 1)), n = 0;
    for (int i = 0; i < n_org && n < count; i++) {
        int slot = (int)(((long long)start + (long long)i * stride) % n_org);
        if (w->org_state[slot] == 1) emig[n++] = slot;
    }
    SapioMigBufHeader *H = (SapioMigBufHeader *)buf;
    size_t off = (sizeof(SapioMigBufHeader) + 15) & ~(size_t)15;
    int k = 0;
    for (; k < n; k++) {
        SapioMigHeader h; sapio_mig_header(w, emig[k], &h);
        h.bytes = (int32_t)sapio_mig_record_bytes(&T, &h);
        if (off + (size_t)h.bytes > cap) break;
        H->offset[k] = (int32_t)off;
        char *rec = buf + off;
        memset(rec, 0, (size_t)h.bytes);
        *(SapioMigHeader *)rec = h;
        uint32_t o = sizeof(SapioMigHeader);
        for (int f = 0; f < T.n; f++) {
            copy_field(1, &T.f[f], &h, T.in_cap, T.f[f].base + (size_t)emig[k] * T.f[f].full_bytes, rec + o);
            o += (sapio_mig_used(&T.f[f], &h) + 15u) & ~15u;
        }
        off += (size_t)h.bytes;
    }
    H->magic = SAPIO_MIG_MAGIC; H->n_records = k; H->used_bytes = (int32_t)off; H->chosen = n; H->stayed = n - k; H->src_tick = (int32_t)w->g_tick[0];
    H->pad[0] = H->pad[1] = 0;
    const int na = w->xa_off[w->p.n_org * MAXC + w->p.n_dead];
    for (int q = 0; q < k; q++) {
        int slot = emig[q];
        for (int c = 0; c < MAXC; c++) {
            int row = slot * MAXC + c;
            w->c_alive[row] = 0;
            for (int e = w->xa_off[row]; e < w->xa_off[row + 1] && e < na; e++) w->xa_other[e] = 0x7FFFFFFFu;
        }
        w->org_state[slot] = 0; w->org_req[slot] = 0; w->org_ic_n[slot] = 0; w->org_ncells[slot] = 0;
    }
    scrub(w);
    free(emig);
    return k;
}

static int rects_block(const double d[4], const double r[4]) {        /* sprites.py:180-182 */
    int inx = (d[0] > r[0] && d[0] < r[2]) || (d[2] > r[0] && d[2] < r[2]) || (r[0] > d[0] && r[2] < d[2]);
    if (!inx) return 0;
    return (d[1] > r[1] && d[1] < r[3]) || (d[3] > r[1] && d[3] < r[3]) || (r[1] > d[1] && r[3] < d[3]);
}
static uint32_t field_offset(const SapioMigTable *T, const SapioMigHeader *h, int idx) {
    uint32_t off = sizeof(SapioMigHeader);
    for (int k = 0; k < idx; k++) off += (sapio_mig_used(&T->f[k], h) + 15u) & ~15u;
    return off;
}

/* places and scatters every record of the buffer; out2 = {accepted, dropped}; returns 0 or < 0 */
int oracle_migrate_in(SapioWorld *w, void *buf_, size_t cap, int32_t *out2) {
    char *buf = (char *)buf_;
    (void)cap;
    const SapioMigBufHeader *H = (const SapioMigBufHeader *)buf;
    int n = H->magic == SAPIO_MIG_MAGIC ? H->n_records : 0;
    if (n > SAPIO_MIG_MAX_RECORDS) n = SAPIO_MIG_MAX_RECORDS;
    SapioMigTable T; sapio_mig_table(&w->p, w, &T);
    const int n_org = w->p.n_org;
    double *res = (double *)malloc(sizeof(double) * 4 * (size_t)n_org);       /* residents' boxes at arrival time */
    for (int o = 0; o < n_org; o++) if (w->org_state[o] == 1) ow_org_rect(w, o, res + 4 * o);
    double *acc = (double *)malloc(sizeof(double) * 4 * (size_t)(n > 0 ? n : 1));
    int n_acc = 0, n_drop = 0, cursor = 0;
    const int32_t dtick = (int32_t)w->g_tick[0] - H->src_tick;
    const int64_t uid0 = w->g_next_uid[0];
    uint8_t *resident = (uint8_t *)malloc((size_t)n_org);
    for (int o = 0; o < n_org; o++) resident[o] = w->org_state[o] == 1;
    for (int r = 0; r < n; r++) {
        char *rec = buf + H->offset[r];
        const SapioMigHeader h = *(const SapioMigHeader *)rec;
        const float *pos = (const float *)(rec + field_offset(&T, &h, T.i_cpos));
        const float *size = (const float *)(rec + field_offset(&T, &h, T.i_csize));
        const uint8_t *alive = (const uint8_t *)(rec + field_offset(&T, &h, T.i_calive));
        double minx = pos[0], maxx = minx, miny = pos[1], maxy = miny;
        for (int k = 0; k < h.ncells; k++) {
            if (alive[k] != SAPIO_CELL_ALIVE) continue;
            double x = pos[2 * k], y = pos[2 * k + 1], rad = (double)size[k] / 2.0;
            if (x - rad < minx) minx = x - rad;
            if (x + rad > maxx) maxx = x + rad;
            if (y - rad < miny) miny = y - rad;
            if (y + rad > maxy) maxy = y + rad;
        }
        const double hx = pos[0], hy = pos[1], R[4] = {minx - hx, miny - hy, maxx - hx, maxy - hy};
        int chosen = -1; double cx = 0, cy = 0;
        for (int k = 0; k < SAPIO_MIG_CANDIDATES && chosen < 0; k++) {
            uint32_t q[4];
            sapio_rng4(w->p.seed, uid0 + r, (int32_t)w->g_tick[0], SAPIO_RNG_MIGRATE, (uint32_t)k, q);
            double x = (double)sapio_u01(q[0]) * (double)w->p.width, y = (double)sapio_u01(q[1]) * (double)w->p.height;
            double d[4] = {R[0] + x, R[1] + y, R[2] + x, R[3] + y};
            if (d[0] < 0 || d[2] > (double)w->p.width || d[1] < 0 || d[3] > (double)w->p.height) continue;
            int blocked = 0;
            for (int o = 0; o < n_org && !blocked; o++) if (resident[o] && rects_block(d, res + 4 * o)) blocked = 1;
            for (int a = 0; a < n_acc && !blocked; a++) if (rects_block(d, acc + 4 * a)) blocked = 1;
            if (!blocked) { chosen = k; cx = x; cy = y; }
        }
        int slot = -1;
        if (chosen >= 0) { while (cursor < n_org && w->org_state[cursor] != 0) cursor++; if (cursor < n_org) slot = cursor++; }
        if (slot < 0) { n_drop++; continue; }
        uint32_t o = sizeof(SapioMigHeader);
        for (int f = 0; f < T.n; f++) {
            copy_field(0, &T.f[f], &h, T.in_cap, T.f[f].base + (size_t)slot * T.f[f].full_bytes, rec + o);
            o += (sapio_mig_used(&T.f[f], &h) + 15u) & ~15u;
        }
        const double dx = cx - hx, dy = cy - hy;
        for (int k = 0; k < MAXC; k++) {
            int row = slot * MAXC + k;
            if (k >= h.ncells) { w->c_alive[row] = 0; continue; }
            w->c_pos[2 * row] = (float)((double)w->c_pos[2 * row] + dx); w->c_pos[2 * row + 1] = (float)((double)w->c_pos[2 * row + 1] + dy);
            w->c_spos[2 * row] = (float)((double)w->c_spos[2 * row] + dx); w->c_spos[2 * row + 1] = (float)((double)w->c_spos[2 * row + 1] + dy);
        }
        w->org_pos[2 * slot] = (float)((double)w->org_pos[2 * slot] + dx); w->org_pos[2 * slot + 1] = (float)((double)w->org_pos[2 * slot + 1] + dy);
        w->org_uid[slot] = uid0 + n_acc;
        w->org_birth_tick[slot] += dtick; w->org_lastbirth_tick[slot] += dtick; w->org_think_tick[slot] += dtick;
        w->org_train_tick[slot] += dtick; w->org_dopa_tick[slot] += dtick;
        w->org_req[slot] = 0;
        double *A = acc + 4 * n_acc;
        A[0] = R[0] + cx; A[1] = R[1] + cy; A[2] = R[2] + cx; A[3] = R[3] + cy;
        n_acc++;
    }
    for (int i = 0; i < w->p.n_dead; i++) {
        if (!w->d_state[i]) continue;
        double x = w->d_pos[2 * i], y = w->d_pos[2 * i + 1];
        for (int a = 0; a < n_acc; a++) {
            const double *d = acc + 4 * a;
            if (x > d[0] && x < d[2] && y > d[1] && y < d[3]) { w->d_state[i] = 0; w->g_stats[SAPIO_STAT_DEAD_REMOVED]++; break; }
        }
    }
    w->g_next_uid[0] += n_acc;
    if (out2) { out2[0] = n_acc; out2[1] = n_drop; }
    free(res); free(acc); free(resident);
    return 0;
}
