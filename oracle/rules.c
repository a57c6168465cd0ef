/* rules.c -- CPU oracle: sprites.update_organisms (sprites.py:1809-1856) and what it drives:
 * Organism.update (:1627-1681), _update_cell (:1470-1606), add/take_energy (:1380-1407), convert_co2/o2 (:1409-1439),
 * kill_cell (:1326-1354), rotation (:1179-1201), _update_brain (:1608-1625), gradual_dispersion (:1752-1772), bounds
 * (:1849-1856), and updateUI's simulation-relevant part (Sapiogenesis.py:57-65).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Canonical (order-free) semantics -- where the reference's result depends on the position of an organism in
 * organism_list, the contract is the one a parallel machine can reproduce (DESIGN.md "Canonical tick"):
 *   C1 every read of ANOTHER organism's cell sees its state at the start of the rules phase (alive flags, positions);
 *   C2 carnivore damage to other organisms' cells and mass eaten from dead cells are summed per victim in attacker-id order
 *      and applied after all organisms have updated (oracle_settle) -- i.e. as if every victim preceded its attackers in
 *      organism_list (sprites.py:1497-1516 applies them immediately);
 *   C3 `view` and `colliding` are the geometric sets of the last Space.step's detection (SURVEY C-5), without the
 *      one-tick-stale entries the callback bookkeeping leaves behind (sprites.py:1481-1489);
 *   C4 the global gas scalars are threaded through organisms in slot order and cells in row order exactly as the
 *      sequential reference does, except that the co2 refund of an energy-saturated organism (sprites.py:1423-1428) is
 *      returned at the end of the rules phase instead of immediately (so each cell's map is linear; SURVEY C-6);
 *   C5 virtual clock: now = g_tick, uDiff = dt_wall for every rule.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

int oracle_think(SapioWorld *w, int o, const double *health, const double *energy, const uint8_t *alive, double rotation_deg);

typedef struct {
    int o, nc;
    double health[MAXC], energy[MAXC], vx[MAXC], vy[MAXC], av[MAXC];
    uint8_t alive[MAXC];      /* 1 alive, 2 died this tick, 0 absent */
    uint8_t inuse[MAXC];
    double consumed;
} Org;

static void org_load(const SapioWorld *w, int o, Org *g) {
    g->o = o; g->nc = w->org_ncells[o];
    for (int k = 0; k < g->nc; k++) {
        int row = o * MAXC + k;
        g->health[k] = w->c_health[row]; g->energy[k] = w->c_energy[row];
        g->vx[k] = w->c_vel[2 * row]; g->vy[k] = w->c_vel[2 * row + 1]; g->av[k] = w->c_angvel[row];
        g->alive[k] = w->c_alive[row]; g->inuse[k] = w->c_inuse[row];
    }
    g->consumed = w->org_consumed[o];
}
static void org_store(SapioWorld *w, const Org *g) {
    int o = g->o;
    for (int k = 0; k < g->nc; k++) {
        int row = o * MAXC + k;
        w->c_health[row] = (float)g->health[k]; w->c_energy[row] = (float)g->energy[k];
        w->c_vel[2 * row] = (float)g->vx[k]; w->c_vel[2 * row + 1] = (float)g->vy[k]; w->c_angvel[row] = (float)g->av[k];
        w->c_alive[row] = g->alive[k]; w->c_inuse[row] = g->inuse[k];
    }
    w->org_consumed[o] = (float)g->consumed;
}

static int n_living(const Org *g) { int n = 0; for (int k = 0; k < g->nc; k++) n += (g->alive[k] == SAPIO_CELL_ALIVE); return n; }
/* Organism.max_energy / total_energy / energy_percent (sprites.py:1306-1315) */
static double max_energy(const SapioWorld *w, const Org *g) { double s = 0; for (int k = 0; k < g->nc; k++) if (g->alive[k] == 1) s += w->c_storage[g->o * MAXC + k]; return s; }
static double total_energy(const Org *g) { double s = 0; for (int k = 0; k < g->nc; k++) if (g->alive[k] == 1) s += g->energy[k]; return s; }
static double energy_percent(const SapioWorld *w, const Org *g) {
    double t = total_energy(g), m = max_energy(w, g);
    if (t == 0 || m == 0) return 0.0;
    return t / m * 100.0;
}
/* Organism.max_health (all created cells) / current_health (living) / health_percent (sprites.py:1317-1324) */
static double health_percent(const SapioWorld *w, const Org *g) {
    double cur = 0, mx = 0;
    for (int k = 0; k < g->nc; k++) { mx += w->c_maxhealth[g->o * MAXC + k]; if (g->alive[k] == 1) cur += g->health[k]; }
    if (cur == 0 || mx == 0) return 0.0;
    return cur / mx * 100.0;
}
/* Organism.add_energy (sprites.py:1380-1398): equal shares in row order, overflow carried to the next living cell,
 * last overflow returned */
static double add_energy(const SapioWorld *w, Org *g, double amount) {
    int n = n_living(g);
    if (!n) return 0.0;
    double per = amount / n, spare = 0;
    for (int k = 0; k < g->nc; k++) {
        if (g->alive[k] != 1) continue;
        double cap = w->c_storage[g->o * MAXC + k];
        g->energy[k] += per + spare;
        spare = 0;
        if (g->energy[k] > cap) { spare = g->energy[k] - cap; g->energy[k] = cap; }
    }
    return spare;
}
/* Organism.take_energy (sprites.py:1400-1407) */
static void take_energy(Org *g, double amount) {
    int n = n_living(g);
    for (int k = 0; k < g->nc; k++) {
        if (g->alive[k] != 1) continue;
        g->energy[k] -= amount / n;
        if (g->energy[k] < 0) g->energy[k] = 0;
    }
}
/* Organism.kill_cell (sprites.py:1326-1354). The dead body itself is created in oracle_settle from this row. */
static void kill_cell(SapioWorld *w, Org *g, int k) {
    if (g->alive[k] != 1) return;
    int row = g->o * MAXC + k;
    g->alive[k] = SAPIO_CELL_DYING;
    g->health[k] = 0;
    int parent = w->c_parent[row];
    if (parent >= 0 && g->alive[parent] == 1)
        g->health[parent] -= (double)w->p.break_damage * (double)w->c_size[row];
    w->g_stats[SAPIO_STAT_CELL_DEATHS]++;
    for (int c = 0; c < g->nc; c++)
        if (c != k && w->c_parent[g->o * MAXC + c] == k) kill_cell(w, g, c);
}
/* Organism.rotation (sprites.py:1179-1201), degrees in [0, 360) */
static double org_rotation(const SapioWorld *w, const Org *g) {
    int rel = -1;
    for (int k = 1; k < g->nc; k++) if (g->alive[k] == 1) { rel = k; break; }
    if (rel < 0) return 0.0;
    int r0 = g->o * MAXC, rr = r0 + rel;
    double old_a = atan2((double)w->c_rel[2 * rr + 1], (double)w->c_rel[2 * rr]);
    double cur_a = atan2((double)w->c_pos[2 * rr + 1] - (double)w->c_pos[2 * r0 + 1], (double)w->c_pos[2 * rr] - (double)w->c_pos[2 * r0]);
    double deg = (cur_a - old_a) * (180.0 / M_PI);
    double m = fmod(deg, 360.0);
    if (m < 0) m += 360.0;
    return m;
}

/* is body `b` (cell row or dead slot id) something a cell of organism `o` registers in `colliding`
 * (Organism._cell_collision, sprites.py:1373-1378)? returns 1 live cell of another organism, 2 dead cell, 0 neither */
static int collider_kind(const SapioWorld *w, int o, uint32_t b) {
    uint32_t NC = (uint32_t)w->p.n_org * MAXC;
    if (b < NC) return ((int)(b / MAXC) != o && w->c_alive[b] != 0) ? 1 : 0;
    if (b < NC + (uint32_t)w->p.n_dead) return w->d_state[b - NC] ? 2 : 0;
    return 0;
}

static void update_cell(SapioWorld *w, Org *g, int k, double dt, double *co2, double *o2, double *refund) {
    const SapioParams *p = &w->p;
    int o = g->o, row = o * MAXC + k;
    int immortal = w->org_flags[o] & SAPIO_F_IMMORTAL;
    if (g->health[k] <= 0 && !immortal) { kill_cell(w, g, k); return; }   /* sprites.py:1474-1476 */
    if (g->alive[k] != 1) return;
    int type = w->c_type[row];
    const float *mv = w->org_move + 4 * o;                                   /* rotation, speed, dir_x, dir_y */

    if (type == SAPIO_T_CARNIV) {                                            /* sprites.py:1491-1516, contract C2/C3 */
        int any = 0;
        double dmg = w->c_damage[row];
        for (int c = 0; c < w->g_x_n[0]; c++) {
            uint32_t a = (uint32_t)(w->x_key[c] >> 32), b = (uint32_t)(w->x_key[c] & 0xFFFFFFFFu), other;
            if (a == (uint32_t)row) other = b; else if (b == (uint32_t)row) other = a; else continue;
            int kind = collider_kind(w, o, other);
            if (!kind) continue;
            any = 1;
            double added = dmg / 2;
            if (kind == 2) { int d = (int)(other - (uint32_t)p->n_org * MAXC); added *= w->d_mass0[d]; added *= w->d_size[d]; }
            g->consumed += added;
        }
        g->inuse[k] = (uint8_t)any;
    }
    if (type == SAPIO_T_PUSH) {                                              /* sprites.py:1519-1544 */
        if (fabs((double)mv[1]) >= 0.1) {
            double ang = atan2((double)mv[3], (double)mv[2]) + org_rotation(w, g) * (M_PI / 180.0);
            double speed = (double)mv[1] * p->max_push * dt;
            g->vx[k] += cos(ang) * speed; g->vy[k] += sin(ang) * speed;
            g->inuse[k] = 1;
        } else g->inuse[k] = 0;
    }
    if (type == SAPIO_T_ROTATE) {                                            /* sprites.py:1546-1552: per tick, not per second */
        g->av[k] += (double)mv[0] * p->max_rot;
        g->inuse[k] = fabs((double)mv[0]) >= 0.1;
    }
    double idle = w->c_use_idle[row], act = w->c_use_act[row], usage;      /* sprites.py:1554-1574 */
    if (g->inuse[k]) {
        if (type == SAPIO_T_PUSH || type == SAPIO_T_ROTATE) {
            double e = (act - idle) * (double)(type == SAPIO_T_PUSH ? mv[1] : mv[0]);
            usage = e < 0 ? fabs(-idle + e) : idle + e;
        } else usage = act;
    } else usage = idle;
    take_energy(g, usage * dt);
    if (g->energy[k] < 0) g->energy[k] = 0;

    double rho = (double)w->c_size[row] / ((double)p->width + (double)p->height) * 100.0 * dt / 100.0;  /* num2perc -> *uDiff -> perc2num */
    if (type == SAPIO_T_CO2C) {                                              /* convert_co2 (sprites.py:1409-1428), contract C4 */
        double gq = rho * (*co2);
        *co2 -= gq; *o2 += gq;
        double spare = add_energy(w, g, gq * p->co2_ratio);
        *refund += spare / p->co2_ratio;
    } else {                                                                 /* convert_o2 (sprites.py:1430-1439) */
        double gq = rho * (*o2);
        *o2 -= gq; *co2 += gq;
    }
    double cap = w->c_storage[row], mh = w->c_maxhealth[row];
    if (g->energy[k] > cap) g->energy[k] = cap;
    if (g->energy[k] >= cap / 2 && g->health[k] < mh) {                      /* heal (sprites.py:1589-1594) */
        g->health[k] += (double)p->heal_rate / 100.0 * mh * dt;
        if (g->health[k] > mh) g->health[k] = mh;
    }
    if (g->energy[k] <= 2.0 / 100.0 * cap)                                   /* starve (sprites.py:1603-1606) */
        g->health[k] -= (double)p->starvation_rate / 100.0 * mh * dt;
}

/* Organism.update (sprites.py:1627-1681) */
static void org_update(SapioWorld *w, int o, uint32_t phases, double *co2, double *o2) {
    const SapioParams *p = &w->p;
    const int32_t T = (int32_t)w->g_tick[0];
    const double dt = p->dt_wall;
    Org g; org_load(w, o, &g);
    w->org_req[o] = 0;
    w->org_gas_in[o] = *co2;
    double refund = 0;
    int immortal = w->org_flags[o] & SAPIO_F_IMMORTAL;

    if (T - w->org_birth_tick[o] >= p->age_ticks && !immortal) kill_cell(w, &g, 0);     /* sprites.py:1826-1827 */
    if (n_living(&g) > 0) {                                                             /* org.alive() (sprites.py:1829) */
        double emax = max_energy(w, &g);                                                /* digestion (sprites.py:1631-1640) */
        if (g.consumed >= emax / 2) g.consumed = emax / 2;
        if (g.consumed < 0.001) g.consumed = 0;
        if (g.consumed != 0) {
            double dig = 95.0 / 100.0 * g.consumed * dt;
            g.consumed -= dig;
            g.consumed += add_energy(w, &g, dig * 1.2);
        }
        if ((phases & SAPIO_PH_THINK) && w->org_nlayers[o] > 0 && T - w->org_think_tick[o] >= p->think_ticks) {
            /* _update_brain (sprites.py:1608-1625) */
            double ddt = (double)(T - w->org_dopa_tick[o]) * dt;
            double hd = health_percent(w, &g) - (double)w->org_last_health[o];
            double ed = energy_percent(w, &g) - (double)w->org_last_energy[o];
            double pain = hd / ddt * 5.0;
            if (pain > 0) pain = 0.0;
            pain = fabs(pain);
            w->org_energy_diff[o] = (float)(ed / ddt);
            w->org_pain[o] = (float)pain;
            w->org_dopamine[o] = (float)(ed / ddt - pain);
            w->org_dopa_tick[o] = T;
            w->org_think_tick[o] = T;
            oracle_think(w, o, g.health, g.energy, g.alive, org_rotation(w, &g));
            w->org_req[o] |= SAPIO_REQ_THOUGHT;
            w->g_stats[SAPIO_STAT_THINKS]++;
        }
        w->org_last_health[o] = (float)health_percent(w, &g);                           /* sprites.py:1653-1654 */
        w->org_last_energy[o] = (float)energy_percent(w, &g);
        if (n_living(&g) == 1) kill_cell(w, &g, 0);                                     /* only the heart left (sprites.py:1656-1657) */
        if (g.alive[0] == 1) { w->org_pos[2 * o] = w->c_pos[2 * (o * MAXC)]; w->org_pos[2 * o + 1] = w->c_pos[2 * (o * MAXC) + 1]; }
        emax = max_energy(w, &g);
        if (n_living(&g) > 0 && total_energy(&g) >= emax / 1.01 && T - w->org_lastbirth_tick[o] >= p->repro_ticks) {
            uint32_t r[4];
            sapio_rng4(p->seed, w->org_uid[o], T, SAPIO_RNG_REPRO_GATE, 0, r);
            double u = sapio_u01(r[0]);
            if (u * 100.0 <= health_percent(w, &g) && !(w->g_population[0] >= p->population_limit)) {
                if (phases & SAPIO_PH_REPRODUCE) w->org_req[o] |= SAPIO_REQ_REPRODUCE;
                w->org_lastbirth_tick[o] = T;                                           /* _reproduce_organism (sprites.py:1442) */
            }
            for (int k = 0; k < g.nc; k++) if (g.alive[k] == 1) g.energy[k] = g.energy[k] / 2;   /* sprites.py:1672-1675 */
            w->org_last_health[o] = (float)health_percent(w, &g);
            w->org_last_energy[o] = (float)energy_percent(w, &g);
        }
        for (int k = 0; k < g.nc; k++) update_cell(w, &g, k, dt, co2, o2, &refund);    /* sprites.py:1680-1681 */
        if ((phases & SAPIO_PH_TRAIN) && !(w->org_flags[o] & SAPIO_F_MALADAPTIVE) && w->org_nlayers[o] > 0 &&
            T - w->org_train_tick[o] >= p->train_ticks)
            w->org_req[o] |= SAPIO_REQ_TRAIN;                                           /* sprites.py:1832-1835 */
        w->g_stats[SAPIO_STAT_ORG_TICKS]++;
    }
    w->org_gas_refund[o] = refund;
    org_store(w, &g);
}

void oracle_view_index_build(const SapioWorld *w);
void oracle_view_index_drop(void);
int oracle_rules(SapioWorld *w, uint32_t phases) {
    double co2 = w->g_co2[0], o2 = w->g_o2[0], refund = 0;
    if (phases & SAPIO_PH_THINK) oracle_view_index_build(w);                   /* candidates for this phase's sensors (brain.c) */
    for (int o = 0; o < w->p.n_org; o++) {
        if (w->org_state[o] != 1) continue;
        org_update(w, o, phases, &co2, &o2);
        refund += w->org_gas_refund[o];
    }
    oracle_view_index_drop();
    co2 += refund; o2 -= refund;                                              /* contract C4: refunds at the end of the phase */
    w->g_co2[0] = co2; w->g_o2[0] = o2;
    return 0;
}

/* After all organisms: apply the deferred carnivore effects (C2), turn cells that died into free dead bodies
 * (_make_dead_cell, sprites.py:1164-1177,1341-1349), gradual_dispersion (sprites.py:1752-1772), bounds (sprites.py:1849-1856),
 * and drop organisms with no living cell (sprites.py:1837). */
int oracle_settle(SapioWorld *w) {
    const SapioParams *p = &w->p;
    const uint32_t NC = (uint32_t)p->n_org * MAXC, ND = (uint32_t)p->n_dead;
    const double dt = p->dt_wall;
    int rc = 0;
    /* C2: gather per victim, attackers in ascending body id. The contact list is sorted by (a, b): for a victim v the partners with a
     * smaller id are the `a` of the contacts with b == v, in list order, followed by the `b` of the contacts with a == v, in list order.
     * Both runs are laid out per victim first (counting sort over the contacts), so the sums run over the same terms in the same order
     * as a scan of the whole list per victim would -- without its bodies x contacts cost. */
    {
        const int nx = w->g_x_n[0];
        const uint32_t NBody = NC + ND;
        int *off = (int *)calloc((size_t)NBody + 2, sizeof(int));
        uint32_t *att = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(2 * (nx > 0 ? nx : 1)));
        /* slots per victim: first its smaller partners, then its larger ones */
        for (int c = 0; c < nx; c++) {
            uint32_t a = (uint32_t)(w->x_key[c] >> 32), b = (uint32_t)(w->x_key[c] & 0xFFFFFFFFu);
            if (b < NBody) off[b + 1]++;
            if (a < NBody) off[a + 1]++;
        }
        for (uint32_t v = 0; v < NBody; v++) off[v + 1] += off[v];
        int *fill = (int *)malloc(sizeof(int) * ((size_t)NBody + 1));
        memcpy(fill, off, sizeof(int) * ((size_t)NBody + 1));
        for (int c = 0; c < nx; c++) {                                                   /* smaller partners, in list order */
            uint32_t a = (uint32_t)(w->x_key[c] >> 32), b = (uint32_t)(w->x_key[c] & 0xFFFFFFFFu);
            if (b < NBody) att[fill[b]++] = a;
        }
        for (int c = 0; c < nx; c++) {                                                   /* then the larger ones, in list order */
            uint32_t a = (uint32_t)(w->x_key[c] >> 32), b = (uint32_t)(w->x_key[c] & 0xFFFFFFFFu);
            if (a < NBody) att[fill[a]++] = b;
        }
        free(fill);
        for (uint32_t v = 0; v < NBody; v++) {
            if (off[v] == off[v + 1]) continue;
            if (v < NC) { if (w->c_alive[v] != SAPIO_CELL_ALIVE) continue; }
            else if (!w->d_state[v - NC]) continue;
            double sum = 0; int hit = 0;
            for (int q = off[v]; q < off[v + 1]; q++) {
                const uint32_t at = att[q];
                if (at >= NC || w->c_alive[at] != SAPIO_CELL_ALIVE || w->c_type[at] != SAPIO_T_CARNIV) continue;
                if (v < NC && at / MAXC == v / MAXC) continue;
                if (v < NC) sum += (double)w->c_damage[at];
                else sum += (double)w->c_damage[at] / (double)w->c_size[at];             /* sprites.py:1505 */
                hit = 1;
            }
            if (!hit) continue;
            if (v < NC) { if (w->c_type[v] != SAPIO_T_BARRIER) w->c_health[v] = (float)((double)w->c_health[v] - sum); }
            else { double m = (double)w->d_mass[v - NC] - sum; w->d_mass[v - NC] = (float)(m > 0 ? m : 0); }
        }
        free(off); free(att);
    }
    /* deaths -> dead pool: the j-th dying cell in (organism, row) order takes the j-th smallest free slot */
    uint32_t next_free = 0;
    for (int o = 0; o < p->n_org; o++) {
        if (w->org_state[o] != 1) continue;
        int living = 0;
        for (int k = 0; k < w->org_ncells[o]; k++) {
            int row = o * MAXC + k;
            if (w->c_alive[row] == SAPIO_CELL_ALIVE) living++;
            if (w->c_alive[row] != SAPIO_CELL_DYING) continue;
            while (next_free < ND && w->d_state[next_free]) next_free++;
            if (next_free >= ND) { w->g_stats[SAPIO_STAT_OVERFLOW]++; rc = SAPIO_E_OVERFLOW; w->c_alive[row] = 0; continue; }
            uint32_t d = next_free++;
            w->d_state[d] = 2;                                                           /* created this tick */
            w->d_pos[2 * d] = w->c_pos[2 * row]; w->d_pos[2 * d + 1] = w->c_pos[2 * row + 1];
            w->d_vel[2 * d] = w->c_vel[2 * row]; w->d_vel[2 * d + 1] = w->c_vel[2 * row + 1];   /* sprites.py:1348-1349 */
            w->d_ang[d] = 0; w->d_angvel[d] = w->c_angvel[row];
            w->d_vbias[2 * d] = w->d_vbias[2 * d + 1] = 0; w->d_wbias[d] = 0;
            w->d_mass[d] = w->d_mass0[d] = w->c_mass[row];
            w->d_size[d] = w->c_size[row]; w->d_elast[d] = w->c_elast[row]; w->d_fric[d] = w->c_fric[row];
            w->d_owner[d] = w->org_uid[o];
            w->c_alive[row] = 0;
            w->c_health[row] = 0;
        }
        if (!living) { w->org_state[o] = 0; w->g_stats[SAPIO_STAT_DEATHS]++; }
    }
    /* gradual_dispersion over dead cells in slot order; the gas exchange is sequential like the reference's sprite loop */
    double co2 = w->g_co2[0], o2 = w->g_o2[0];
    for (uint32_t d = 0; d < ND; d++) {
        if (!w->d_state[d]) continue;
        double mass = w->d_mass[d];
        int remove = 0;
        if (mass < p->mass_cutoff) remove = 1;
        else {
            double rate = (double)p->dispersion_rate / 100.0 * mass * dt;
            double nm = mass - rate;
            if (nm > 0) {
                double g = rate / 100.0 * o2;                                            /* perc2num(dispersion_rate, oxygen): a percentage */
                w->d_mass[d] = (float)nm;
                co2 += g; o2 -= g;
            } else remove = 1;
        }
        double x = w->d_pos[2 * d], y = w->d_pos[2 * d + 1];
        if (x < 0 || x > p->width || y < 0 || y > p->height) remove = 1;               /* sprites.py:1850-1856 */
        if (remove) { w->d_state[d] = 0; w->g_stats[SAPIO_STAT_DEAD_REMOVED]++; }
    }
    w->g_co2[0] = co2; w->g_o2[0] = o2;
    /* living cells out of bounds: health = 0 (they die at their next _update_cell) */
    for (int o = 0; o < p->n_org; o++) {
        if (w->org_state[o] != 1) continue;
        for (int k = 0; k < w->org_ncells[o]; k++) {
            int row = o * MAXC + k;
            if (w->c_alive[row] != SAPIO_CELL_ALIVE) continue;
            double x = w->c_pos[2 * row], y = w->c_pos[2 * row + 1];
            if (x < 0 || x > p->width || y < 0 || y > p->height) w->c_health[row] = 0;
        }
    }
    return rc;
}

/* updateUI (Sapiogenesis.py:57-65): clamp the gases at 0, recount the population */
int oracle_post(SapioWorld *w) {
    if (w->g_co2[0] < 0) w->g_co2[0] = 0;
    if (w->g_o2[0] < 0) w->g_o2[0] = 0;
    int pop = 0;
    for (int o = 0; o < w->p.n_org; o++) pop += (w->org_state[o] == 1);
    w->g_population[0] = pop;
    return 0;
}

/* test hook: neural.activate for organism o on the world's current state (dopamine/pain as stored) */
int oracle_think_now(SapioWorld *w, int o) {
    Org g; org_load(w, o, &g);
    return oracle_think(w, o, g.health, g.energy, g.alive, org_rotation(w, &g));
}
double oracle_rotation(const SapioWorld *w, int o) {
    Org g; org_load(w, o, &g);
    return org_rotation(w, &g);
}

/* updateUI world events (Sapiogenesis.py:37-53): drought takes `drought` from both gases, poison takes 5 * `poison` health
 * from every cell of every organism, with probability `random_cells` a piece of food appears; all = severity * uDiff, computed by the host */
int oracle_events(SapioWorld *w, double drought, double poison, double random_cells) {
    if (poison != 0)
        for (int o = 0; o < w->p.n_org; o++) {
            if (w->org_state[o] != 1) continue;
            for (int k = 0; k < w->org_ncells[o]; k++) w->c_health[o * MAXC + k] = (float)((double)w->c_health[o * MAXC + k] - poison * 5.0);
        }
    if (drought != 0) { w->g_o2[0] -= drought; w->g_co2[0] -= drought; }
    if (random_cells > 0) {   /* Sapiogenesis.py:48-53: random.random() <= rVal -> add_dead_clicked at (randrange(10, int(W)), randrange(10, int(H))) */
        uint32_t r[4];
        sapio_rng4(w->p.seed, -1, (int32_t)w->g_tick[0], 11u, 0u, r);
        if (sapio_u01(r[0]) <= (float)random_cells) {
            int d = 0;
            while (d < w->p.n_dead && w->d_state[d] != 0) d++;
            if (d >= w->p.n_dead) { w->g_stats[SAPIO_STAT_OVERFLOW] += 1; return 0; }
            const int wi = (int)w->p.width, hi = (int)w->p.height;
            w->d_state[d] = 2;
            w->d_pos[2 * d] = (float)(10 + (int)floor((double)sapio_u01(r[1]) * (double)(wi - 10)));
            w->d_pos[2 * d + 1] = (float)(10 + (int)floor((double)sapio_u01(r[2]) * (double)(hi - 10)));
            w->d_vel[2 * d] = w->d_vel[2 * d + 1] = 0.f; w->d_ang[d] = w->d_angvel[d] = 0.f;
            w->d_vbias[2 * d] = w->d_vbias[2 * d + 1] = 0.f; w->d_wbias[d] = 0.f;
            w->d_mass[d] = w->d_mass0[d] = 12.f; w->d_size[d] = 30.f; w->d_elast[d] = w->d_fric[d] = 0.5f; w->d_owner[d] = -1; w->d_eaten[d] = 0.f;
        }
    }
    return 0;
}
