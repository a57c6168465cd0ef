/* reproduce.c -- CPU oracle: Organism.reproduce / _reproduce_organism (sprites.py:1441-1459),
 * get_new_organism_position (sprites.py:194-219), world tick driver (physics.py:425-442) and the entry points the tests load.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Births are processed after every organism has updated (the reference does it inside Organism.update, sprites.py:1663-1670),
 * parents in slot order; a birth sees the newborns accepted before it, as in the reference's list order.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>


static USrc stream_src(const SapioParams *p, int64_t uid, int32_t tick, uint32_t purpose, uint32_t start) {
    USrc u; memset(&u, 0, sizeof(u));
    u.st.seed = p->seed; u.st.uid = uid; u.st.tick = tick; u.st.purpose = purpose; u.st.n = start;
    return u;
}
static USrc array_src(const float *a, int n) { USrc u; memset(&u, 0, sizeof(u)); u.arr = a; u.n_arr = n; return u; }

static int free_org_slot(const SapioWorld *w) {
    for (int o = 0; o < w->p.n_org; o++) if (w->org_state[o] == 0) return o;
    return -1;
}

/* dead cells whose centre lies strictly inside the newborn's rect are deleted (sprites.py:184-190, 215-217) */
static void clear_footprint(SapioWorld *w, const OGenome *g, double px, double py) {
    double d[4];
    og_get_rect(g, d);
    d[0] += px; d[1] += py; d[2] += px; d[3] += py;
    for (int i = 0; i < w->p.n_dead; i++) {
        if (!w->d_state[i]) continue;
        double x = w->d_pos[2 * i], y = w->d_pos[2 * i + 1];
        if (x > d[0] && x < d[2] && y > d[1] && y < d[3]) { w->d_state[i] = 0; w->g_stats[SAPIO_STAT_DEAD_REMOVED]++; }
    }
}

/* get_new_organism_position: the 360 integer bearings in random order (sorted by one uniform each, stable), first viable */
static int new_organism_position(const SapioWorld *w, int parent, const OGenome *pg, const OGenome *cg, const float *keys, double pos[2]) {
    double os[2], ns[2];
    og_get_size(pg, os); og_get_size(cg, ns);
    double dist[2] = { os[0] / 1.3 + ns[0] / 1.3, os[1] / 1.3 + ns[1] / 1.3 };
    int best = -1;
    for (int b = 0; b < 360; b++) {
        if (best >= 0 && !(keys[b] < keys[best])) continue;
        double rad = b * (M_PI / 180.0);
        double x = (double)w->org_pos[2 * parent] + cos(rad) * dist[0];
        double y = (double)w->org_pos[2 * parent + 1] + sin(rad) * dist[1];
        if (ow_viable_position(w, cg, x, y)) { best = b; pos[0] = x; pos[1] = y; }
    }
    return best;
}

int oracle_reproduce(SapioWorld *w) {
    const SapioParams *p = &w->p;
    const int32_t T = (int32_t)w->g_tick[0];
    int rc = 0;
    OGenome pg, cg; og_init(&pg); og_init(&cg);
    for (int o = 0; o < p->n_org; o++) {
        if (w->org_state[o] != 1 || !(w->org_req[o] & SAPIO_REQ_REPRODUCE)) continue;
        w->org_req[o] &= (uint8_t)~SAPIO_REQ_REPRODUCE;
        for (int i = 0; i < p->offspring_amount; i++) {
            ow_genome_from_world(w, o, &pg);
            USrc u = stream_src(p, w->org_uid[o], T, SAPIO_RNG_GENOME, (uint32_t)i << 20);
            USrc uf = stream_src(p, w->org_uid[o], T, SAPIO_RNG_WEIGHT_FILL, (uint32_t)i << 20);
            USrc ug = stream_src(p, w->org_uid[o], T, SAPIO_RNG_WEIGHT_GARBAGE, (uint32_t)i << 20);
            if (og_mutate(&cg, &pg, p, p->mutation_severity, p->brain_mutation_severity, p->weight_persistence, &u, &uf, &ug) < 0 ||
                !og_fits_caps(&cg, p)) { w->g_stats[SAPIO_STAT_BIRTH_FAIL_CAPS]++; break; }
            float keys[360];
            for (int b = 0; b < 360; b += 4) {
                uint32_t r[4];
                sapio_rng4(p->seed, w->org_uid[o], T, SAPIO_RNG_BEARINGS, (uint32_t)(i * 90 + b / 4), r);
                for (int q = 0; q < 4; q++) keys[b + q] = sapio_u01(r[q]);
            }
            double pos[2];
            if (new_organism_position(w, o, &pg, &cg, keys, pos) < 0) { w->g_stats[SAPIO_STAT_BIRTH_FAIL_PLACE]++; break; }
            int slot = free_org_slot(w);
            if (slot < 0) { w->g_stats[SAPIO_STAT_OVERFLOW]++; rc = SAPIO_E_OVERFLOW; break; }
            clear_footprint(w, &cg, pos[0], pos[1]);
            int64_t uid = w->g_next_uid[0]++;
            cg.generation = pg.generation + 1;
            USrc ui = stream_src(p, uid, T, SAPIO_RNG_BRAIN_INIT, 0);
            if (ow_materialize(w, slot, &cg, pos[0], pos[1], uid, T, &ui) < 0) { w->g_stats[SAPIO_STAT_BIRTH_FAIL_CAPS]++; w->org_state[slot] = 0; break; }
            ow_inherit_recurrent(w, o, slot);
            w->g_stats[SAPIO_STAT_BIRTHS]++;
        }
    }
    og_free(&pg); og_free(&cg);
    return rc;
}

/* Environment.update (physics.py:425-442) with the hooks of Sapiogenesis.py:687-688 */
int oracle_tick(SapioWorld *w, uint32_t phases, int n_ticks) {
    int rc = 0, r;
    for (int t = 0; t < n_ticks; t++) {
        if (phases & SAPIO_PH_RULES) {
            if ((r = oracle_rules(w, phases)) < 0) rc = r;
            if (phases & SAPIO_PH_TRAIN) if ((r = oracle_train(w)) < 0) rc = r;
            if ((r = oracle_settle(w)) < 0) rc = r;
            if (phases & SAPIO_PH_REPRODUCE) if ((r = oracle_reproduce(w)) < 0) rc = r;
        }
        if (phases & SAPIO_PH_STEP) if ((r = oracle_space_step(w)) < 0) rc = r;
        if (phases & SAPIO_PH_FRICTION) if ((r = oracle_friction(w)) < 0) rc = r;
        if (phases & SAPIO_PH_POST) if ((r = oracle_post(w)) < 0) rc = r;
        w->g_tick[0]++;
    }
    return rc;
}

/* ---- spawn recipe: add_creature_clicked (Sapiogenesis.py:297-323) = DNA().randomize(widget ranges) + Organism + build_body ---- */
int oracle_spawn(SapioWorld *w, int slot, double px, double py) {
    const SapioParams *p = &w->p;
    const int32_t T = (int32_t)w->g_tick[0];
    OGenome g; og_init(&g);
    int64_t uid = w->g_next_uid[0]++;
    int rc = -1;
    for (int attempt = 0; attempt < 64 && rc < 0; attempt++) {                       /* re-roll genomes that exceed the tile caps (ours) */
        USrc u = stream_src(p, uid, T, SAPIO_RNG_GENOME, (uint32_t)attempt << 18);
        USrc ui = stream_src(p, uid, T, SAPIO_RNG_GENOME_INIT, 0);
        if (og_randomize(&g, p, &u, &ui) < 0 || !og_fits_caps(&g, p)) continue;
        USrc ub = stream_src(p, uid, T, SAPIO_RNG_BRAIN_INIT, 0);
        rc = ow_materialize(w, slot, &g, px, py, uid, T, &ub);
    }
    og_free(&g);
    return rc;
}

/* pinning hooks: same code paths, uniforms supplied by the caller (the tests feed the REFERENCE the same arrays) */
int oracle_spawn_arr(SapioWorld *w, int slot, double px, double py, const float *u, int nu, const float *ui, int nui,
                     const float *ub, int nub, int *consumed) {
    OGenome g; og_init(&g);
    USrc su = array_src(u, nu), si = array_src(ui, nui), sb = array_src(ub, nub);
    int rc = og_randomize(&g, &w->p, &su, &si);
    if (rc == 0 && !og_fits_caps(&g, &w->p)) rc = -2;
    if (rc == 0) rc = ow_materialize(w, slot, &g, px, py, w->g_next_uid[0]++, (int32_t)w->g_tick[0], &sb);
    if (consumed) { consumed[0] = su.pos; consumed[1] = si.pos; consumed[2] = sb.pos; consumed[3] = su.overrun | si.overrun | sb.overrun; }
    og_free(&g);
    return rc;
}
int oracle_mutate_arr(SapioWorld *w, int parent, int slot, double px, double py, const float *u, int nu, const float *uf, int nuf,
                      const float *ub, int nub, int *consumed) {
    OGenome pg, cg; og_init(&pg); og_init(&cg);
    USrc su = array_src(u, nu), sf = array_src(uf, nuf), sb = array_src(ub, nub);
    USrc sg = stream_src(&w->p, 0, 0, SAPIO_RNG_WEIGHT_GARBAGE, 0);
    ow_genome_from_world(w, parent, &pg);
    int rc = og_mutate(&cg, &pg, &w->p, w->p.mutation_severity, w->p.brain_mutation_severity, w->p.weight_persistence, &su, &sf, &sg);
    if (rc == 0 && !og_fits_caps(&cg, &w->p)) rc = -2;
    cg.generation = pg.generation + 1;
    if (rc == 0) rc = ow_materialize(w, slot, &cg, px, py, w->g_next_uid[0]++, (int32_t)w->g_tick[0], &sb);
    if (consumed) { consumed[0] = su.pos; consumed[1] = sf.pos; consumed[2] = sb.pos; consumed[3] = su.overrun | sf.overrun | sb.overrun; }
    og_free(&pg); og_free(&cg);
    return rc;
}
/* DNA.get_rect / get_size of a living organism's genome, and viable_organism_position, for the placement pin */
int oracle_dna_rect(const SapioWorld *w, int org, double rect[4], double size[2]) {
    OGenome g; og_init(&g);
    ow_genome_from_world(w, org, &g);
    og_get_rect(&g, rect); og_get_size(&g, size);
    og_free(&g);
    return 0;
}
int oracle_viable(const SapioWorld *w, int genome_of, double px, double py) {
    OGenome g; og_init(&g);
    ow_genome_from_world(w, genome_of, &g);
    int v = ow_viable_position(w, &g, px, py);
    og_free(&g);
    return v;
}
void oracle_org_rect(const SapioWorld *w, int org, double r[4]) { ow_org_rect(w, org, r); }
void oracle_stream_fill(uint64_t seed, int64_t uid, int32_t tick, uint32_t purpose, uint32_t start, int n, float *out) {
    SapioStream s = { seed, uid, tick, purpose, start };
    for (int i = 0; i < n; i++) out[i] = sapio_stream_u(&s);
}
