/* world.c -- CPU oracle: genome <-> world rows, Organism.build_body, placement predicates.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* H = hidden-state columns the input layer of an RNN brain has on top of ldim[0] (networks.py:116) */
static int brain_layer_offset(const int32_t *ldim, int layer, int H) {
    int off = 0;
    for (int l = 0; l < layer; l++) off += ldim[l + 1] * (ldim[l] + (l == 0 ? H : 0)) + ldim[l + 1];
    return off;
}

/* rebuild the DNA object of a living organism from its rows (dna.cells order via c_gorder) */
void ow_genome_from_world(const SapioWorld *w, int org, OGenome *g) {
    float *keep = g->hw; og_init(g); free(keep);
    int nc = w->org_ncells[org];
    g->n = nc;
    for (int k = 0; k < nc; k++) {
        int row = org * MAXC + k, j = w->c_gorder[row];
        g->id[j] = w->c_id[row];
        g->type[j] = w->c_type[row];
        g->first[j] = (w->c_parent[row] < 0);
        g->parent_id[j] = w->c_parent[row] < 0 ? 0 : w->c_id[org * MAXC + w->c_parent[row]];
        g->mirror_id[j] = w->c_mirror[row] < 0 ? 0 : w->c_id[org * MAXC + w->c_mirror[row]];
        g->size[j] = w->c_size[row]; g->mass[j] = w->c_mass[row];
        g->elast[j] = w->c_elast[row]; g->fric[j] = w->c_fric[row];
        g->relx[j] = w->c_rel[2 * row]; g->rely[j] = w->c_rel[2 * row + 1];
        g->growx[j] = w->c_grow[2 * row]; g->growy[j] = w->c_grow[2 * row + 1];
        g->maxhealth[j] = w->c_maxhealth[row]; g->use_idle[j] = w->c_use_idle[row];
        g->use_act[j] = w->c_use_act[row]; g->storage[j] = w->c_storage[row]; g->damage[j] = w->c_damage[row];
    }
    uint8_t fl = w->org_flags[org];
    g->mirror_x = (fl & SAPIO_F_MIRROR_X) != 0; g->mirror_y = (fl & SAPIO_F_MIRROR_Y) != 0;
    g->curiosity = w->org_curiosity[org];
    g->generation = w->org_generation[org];
    const int32_t *ldim = w->org_ldim + org * 16;
    int nl = w->org_nlayers[org];
    g->n_hidden = nl - 1;
    for (int i = 0; i < g->n_hidden; i++) g->hidden[i] = ldim[i + 1];
    int len = 0;
    for (int l = 1; l + 1 < nl; l++) len += ldim[l + 1] * ldim[l];
    g->hw = (float *)malloc(sizeof(float) * (size_t)(len > 0 ? len : 1));
    g->hw_len = len;
    const float *brain = w->org_brain + (size_t)org * w->p.brain_cap;
    int off = 0;
    for (int l = 1; l + 1 < nl; l++) {
        int n = ldim[l + 1] * ldim[l];
        memcpy(g->hw + off, brain + brain_layer_offset(ldim, l, w->org_rnn_h[org]), sizeof(float) * (size_t)n);
        off += n;
    }
}

/* Organism.__init__ + build_body + build_brain + build_sprite (sprites.py:1089-1125, 1683-1729, 255-325).
 * u_init feeds the fresh nn.Linear init of every layer (networks.py:220-233); hidden weights are then overwritten from
 * the genome (sprites.py:1692-1696). */
int ow_materialize(SapioWorld *w, int org, const OGenome *g, double px, double py, int64_t uid, int32_t tick, USrc *u_init) {
    const SapioParams *p = &w->p;
    int order[MAXC];
    int nc = og_instance_order(g, order);
    if (nc != g->n || !og_fits_caps(g, p)) return -1;
    int row_of[MAXC];
    for (int k = 0; k < nc; k++) row_of[order[k]] = k;

    w->org_state[org] = 1;
    w->org_flags[org] = (uint8_t)(SAPIO_F_FINITE_MEMORY | (g->mirror_x ? SAPIO_F_MIRROR_X : 0) | (g->mirror_y ? SAPIO_F_MIRROR_Y : 0));
    w->org_ncells[org] = nc;
    w->org_uid[org] = uid;
    w->org_generation[org] = g->generation;
    w->org_birth_tick[org] = w->org_lastbirth_tick[org] = w->org_think_tick[org] = tick;
    w->org_train_tick[org] = w->org_dopa_tick[org] = tick;
    w->org_consumed[org] = 0;
    for (int i = 0; i < 4; i++) w->org_move[4 * org + i] = 0;
    w->org_dopamine[org] = w->org_pain[org] = w->org_energy_diff[org] = 0;
    w->org_pos[2 * org] = (float)px; w->org_pos[2 * org + 1] = (float)py;
    w->org_curiosity[org] = (float)g->curiosity;
    w->org_boredom[org] = w->org_stimulation[org] = 0;
    for (int i = 0; i < 10; i++) w->org_stim_hist[10 * org + i] = 0;
    w->org_hist_n[org] = w->org_stm_n[org] = w->org_mem_n[org] = 0;
    w->org_adam_t[org] = 0; w->org_loss[org] = 0;
    for (int i = 0; i < 4; i++) w->org_out_raw[4 * org + i] = 0;
    w->org_ic_n[org] = 0;
    w->org_req[org] = 0;
    w->org_gas_a[org] = 1.0; w->org_gas_b[org] = 0.0; w->org_gas_in[org] = 0.0; w->org_gas_refund[org] = 0.0;

    double sum_h = 0, sum_mh = 0, sum_e = 0, sum_st = 0;
    for (int k = 0; k < MAXC; k++) {
        int row = org * MAXC + k;
        if (k >= nc) { w->c_alive[row] = 0; continue; }
        int j = order[k];
        w->c_alive[row] = SAPIO_CELL_ALIVE;
        w->c_type[row] = (uint8_t)g->type[j];
        w->c_inuse[row] = 0;
        int pj = og_index_of(g, g->parent_id[j]);
        w->c_parent[row] = (int8_t)(g->first[j] || pj < 0 ? -1 : row_of[pj]);
        int mj = g->mirror_id[j] ? og_index_of(g, g->mirror_id[j]) : -1;
        w->c_mirror[row] = (int8_t)(mj < 0 ? -1 : row_of[mj]);
        w->c_gorder[row] = (uint8_t)j;
        w->c_id[row] = (int16_t)g->id[j];
        float x = (float)(px + g->relx[j]), y = (float)(py + g->rely[j]);   /* _growth_position (sprites.py:1127-1133) */
        w->c_pos[2 * row] = x; w->c_pos[2 * row + 1] = y;
        w->c_vel[2 * row] = w->c_vel[2 * row + 1] = 0;
        w->c_ang[row] = w->c_angvel[row] = 0;
        w->c_vbias[2 * row] = w->c_vbias[2 * row + 1] = 0; w->c_wbias[row] = 0;
        w->c_spos[2 * row] = x; w->c_spos[2 * row + 1] = y;                 /* newBody.position = pos (sprites.py:307,319) */
        w->c_svel[2 * row] = w->c_svel[2 * row + 1] = 0;
        w->c_svbias[2 * row] = w->c_svbias[2 * row + 1] = 0;
        w->c_spin_jn[row] = 0;
        w->c_health[row] = (float)g->maxhealth[j];                          /* sprites.py:282 */
        w->c_energy[row] = (float)(g->storage[j] / 2.0);                    /* sprites.py:283 */
        w->c_size[row] = (float)g->size[j]; w->c_mass[row] = (float)g->mass[j];
        w->c_elast[row] = (float)g->elast[j]; w->c_fric[row] = (float)g->fric[j];
        w->c_rel[2 * row] = (float)g->relx[j]; w->c_rel[2 * row + 1] = (float)g->rely[j];
        w->c_grow[2 * row] = (float)g->growx[j]; w->c_grow[2 * row + 1] = (float)g->growy[j];
        w->c_maxhealth[row] = (float)g->maxhealth[j]; w->c_use_idle[row] = (float)g->use_idle[j];
        w->c_use_act[row] = (float)g->use_act[j]; w->c_storage[row] = (float)g->storage[j];
        w->c_damage[row] = (float)g->damage[j];
        w->c_dmg[row] = 0;
        sum_h += w->c_health[row]; sum_mh += w->c_maxhealth[row];
        sum_e += w->c_energy[row]; sum_st += w->c_storage[row];
    }
    for (int q = 0; q < SAPIO_NPAIR; q++) w->j_jn[(size_t)org * SAPIO_NPAIR + q] = 0;
    /* build_body tail (sprites.py:1728-1729) */
    w->org_last_health[org] = (float)(sum_mh > 0 && sum_h != 0 ? sum_h / sum_mh * 100.0 : 0.0);
    w->org_last_energy[org] = (float)(sum_st > 0 && sum_e != 0 ? sum_e / sum_st * 100.0 : 0.0);

    /* neural.setup_network (neural.py:426-492): input cells = eyes then olfactory, each in dna.cells order */
    int nin = 0;
    for (int j = 0; j < g->n; j++) if (g->type[j] == SAPIO_T_EYE) w->org_incell[org * MAXC + nin++] = (uint8_t)row_of[j];
    for (int j = 0; j < g->n; j++) if (g->type[j] == SAPIO_T_OLFACTORY) w->org_incell[org * MAXC + nin++] = (uint8_t)row_of[j];
    w->org_nincell[org] = nin;
    int32_t *ldim = w->org_ldim + org * 16;
    memset(ldim, 0, sizeof(int32_t) * 16);
    int nd = 0;
    ldim[nd++] = og_input_size(g);
    for (int i = 0; i < g->n_hidden; i++) ldim[nd++] = g->hidden[i];
    ldim[nd++] = 4;
    int nl = nd - 1;
    w->org_nlayers[org] = nl;
    /* build_brain (sprites.py:1683-1690): an RNNetwork when Environment.info["use_rnn"] is set now -- the input layer takes
     * hiddenSize more columns and a second head outputLayerH is constructed after the output layer (networks.py:116,151-166) */
    const int H = p->rnn_hidden;
    w->org_rnn_h[org] = H;
    w->org_hm_n[2 * org] = w->org_hm_n[2 * org + 1] = 0; w->org_th_n[org] = 0; w->org_mem_stamp[org] = 0;
    for (int i = 0; i < p->rnn_cap; i++) w->org_last_hidden[(size_t)org * p->rnn_cap + i] = 0;      /* lastHidden = zeros (networks.py:149) */
    float *brain = w->org_brain + (size_t)org * p->brain_cap;
    int off = 0, hoff = 0;
    for (int l = 0; l < nl + (H > 0 ? 1 : 0); l++) {
        int fin = l == nl ? ldim[nl - 1] : ldim[l] + (l == 0 ? H : 0), fout = l == nl ? H : ldim[l + 1];
        double bound = 1.0 / sqrt((double)fin);
        for (int e = 0; e < fout * fin; e++) brain[off + e] = (float)((2.0 * usrc_next(u_init) - 1.0) * bound);
        if (l >= 1 && l + 1 < nl) {                                         /* inherited hidden weights */
            memcpy(brain + off, g->hw + hoff, sizeof(float) * (size_t)(fout * fin));
            hoff += fout * fin;
        }
        off += fout * fin;
        for (int e = 0; e < fout; e++) brain[off + e] = (float)((2.0 * usrc_next(u_init) - 1.0) * bound);
        off += fout;
    }
    size_t ac = (size_t)p->adam_cap;
    memset(w->org_adam_m + (size_t)org * ac, 0, sizeof(float) * ac);
    memset(w->org_adam_v + (size_t)org * ac, 0, sizeof(float) * ac);
    return 0;
}

/* What DNA.mutate's deepcopy carries over and erase_memory (sprites.py:413-421) forgets to clear: hidden_memory and trainingHidden
 * (sprites.py:377,386). The child's hidden_memory therefore starts with the parent's entries and runs out of step with its own
 * short-term memory. Cap (ours): the first SAPIO_HM - SAPIO_STM entries are inherited, so that the list never exceeds SAPIO_HM (the
 * reference's list grows by up to 30 entries per generation without bound). */
void ow_inherit_recurrent(SapioWorld *w, int parent, int child) {
    const int RC = w->p.rnn_cap;
    if (RC <= 0) return;
    const int app = w->org_hm_n[2 * parent], pop = w->org_hm_n[2 * parent + 1];
    int n = app - pop;
    if (n > SAPIO_HM - SAPIO_STM) n = SAPIO_HM - SAPIO_STM;
    const float *ph = w->org_hm + (size_t)parent * SAPIO_HM * RC;
    float *ch = w->org_hm + (size_t)child * SAPIO_HM * RC;
    for (int i = 0; i < n; i++) memcpy(ch + (size_t)i * RC, ph + (size_t)((pop + i) % SAPIO_HM) * RC, sizeof(float) * (size_t)RC);
    w->org_hm_n[2 * child] = n; w->org_hm_n[2 * child + 1] = 0;
    const int tn = w->org_th_n[parent];
    memcpy(w->org_th + (size_t)child * SAPIO_MEMROWS * RC, w->org_th + (size_t)parent * SAPIO_MEMROWS * RC, sizeof(float) * (size_t)tn * RC);
    w->org_th_n[child] = tn;
}

/* Organism.get_rect (sprites.py:1211-1232): seeded with the first cell's centre, widened by living cells' discs */
void ow_org_rect(const SapioWorld *w, int org, double r[4]) {
    int row0 = org * MAXC;
    double minx = w->c_pos[2 * row0], maxx = minx, miny = w->c_pos[2 * row0 + 1], maxy = miny;
    for (int k = 0; k < w->org_ncells[org]; k++) {
        int row = row0 + k;
        if (w->c_alive[row] != SAPIO_CELL_ALIVE) continue;
        double x = w->c_pos[2 * row], y = w->c_pos[2 * row + 1], rad = w->c_size[row] / 2.0;
        if (x - rad < minx) minx = x - rad;
        if (x + rad > maxx) maxx = x + rad;
        if (y - rad < miny) miny = y - rad;
        if (y + rad > maxy) maxy = y + rad;
    }
    r[0] = minx; r[1] = miny; r[2] = maxx; r[3] = maxy;
}

/* viable_organism_position (sprites.py:156-192) without the dead-cell collection (done by the caller) */
int ow_viable_position(const SapioWorld *w, const OGenome *g, double px, double py) {
    double d[4];
    og_get_rect(g, d);
    d[0] += px; d[1] += py; d[2] += px; d[3] += py;
    if (d[0] < 0 || d[2] > w->p.width || d[1] < 0 || d[3] > w->p.height) return 0;
    for (int o = 0; o < w->p.n_org; o++) {
        if (w->org_state[o] != 1) continue;
        double r[4];
        ow_org_rect(w, o, r);
        int inx = (d[0] > r[0] && d[0] < r[2]) || (d[2] > r[0] && d[2] < r[2]) || (r[0] > d[0] && r[2] < d[2]);
        if (!inx) continue;
        int iny = (d[1] > r[1] && d[1] < r[3]) || (d[3] > r[1] && d[3] < r[3]) || (r[1] > d[1] && r[3] < d[3]);
        if (iny) return 0;
    }
    return 1;
}
