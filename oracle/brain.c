/* brain.c -- CPU oracle: neural.activate (neural.py:91-366), calculateStimulation (:60-75), train_network (:369-423),
 * networks.Network.forward (networks.py:301-308), Trainer.train_epoch (networks.py:75-87) with torch.optim.Adam defaults.
 * TEST INFRASTRUCTURE ONLY (see oracle.h). PINNED against the reference's own neural.py / networks.py (tests/golden/).
 * The network is fp32 like torch's; scalar bookkeeping is double like Python's.
 */
#include "oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* H = hidden-state columns of an RNN brain's input layer on top of ldim[0] (networks.py:116); 0 for networks.Network */
static int layer_off(const int32_t *ldim, int layer, int H) {
    int off = 0;
    for (int l = 0; l < layer; l++) off += ldim[l + 1] * (ldim[l] + (l == 0 ? H : 0)) + ldim[l + 1];
    return off;
}

/* Network.forward: out = W_out(...W_1(sigmoid(W_in x + b_in)) + b_1...) + b_out -- sigmoid after the first layer only */
/* RNNetwork.forward (networks.py:160-175) when H > 0: x holds ldim[0] inputs followed by the H hidden values (torch.cat), and the
 * second head outputLayerH -- stored after the output layer -- maps the last hidden activation to the next hidden state `hid`. */
void oracle_forward_rnn(const float *brain, const int32_t *ldim, int nl, int H, const float *x, float *out, float *acts, float *hid) {
    float buf[2][1024];
    const float *cur = x, *last = x;
    int off = 0;
    for (int l = 0; l < nl + (H > 0 ? 1 : 0); l++) {
        int fin = l == nl ? ldim[nl - 1] : ldim[l] + (l == 0 ? H : 0), fout = l == nl ? H : ldim[l + 1];
        const float *W = brain + off, *b = W + fout * fin;
        float *dst = (l == nl - 1) ? out : (l == nl) ? hid : buf[l & 1];
        if (l == nl - 1) last = cur;
        if (l == nl) cur = last;
        for (int r = 0; r < fout; r++) {
            float s = 0.0f;
            for (int c = 0; c < fin; c++) s += W[r * fin + c] * cur[c];
            s += b[r];
            if (l == 0) s = 1.0f / (1.0f + expf(-s));
            dst[r] = s;
        }
        if (acts && l == 0) memcpy(acts, dst, sizeof(float) * (size_t)fout);
        cur = dst;
        off += fout * fin + fout;
    }
}
void oracle_forward(const float *brain, const int32_t *ldim, int nl, const float *x, float *out, float *acts) {
    oracle_forward_rnn(brain, ldim, nl, 0, x, out, acts, NULL);
}

/* Python round(x, 1) of a float, times 10, as an integer key (neural.py:77-78). Correctly rounded, ties decided on the
 * exact binary value (glibc printf does exactly that). */
int oracle_key1(double x) {
    char s[64];
    snprintf(s, sizeof(s), "%.1f", x);
    double r = strtod(s, NULL) * 10.0;
    return (int)lrint(r);
}
/* key of a short-term-memory input: it is a Python float holding the fp32 value (tensor.tolist()) */
static int key_stm(float x) { return oracle_key1((double)x); }
/* key of a stored training input: roundList(mem[0], 3) re-rounded the value to the double nearest k/1000 (neural.py:393) */
static double stored_value(float x) {
    char s[64];
    snprintf(s, sizeof(s), "%.3f", (double)x);
    return strtod(s, NULL);
}
static int key_mem(float x) { return oracle_key1(stored_value(x)); }

/* view of one sensor == overlap of the sensor disc (centre: the SENSOR BODY) with solid cells of other organisms and dead
 * cells not descended from this organism (eye_segment_handler, sprites.py:221-236; contract C3). Calls cb per member in
 * ascending body id. Predicate in double on fp32 positions, strict (SURVEY A-3). */
typedef void (*view_cb)(void *ctx, uint32_t body, double x, double y, double r, int alive);

/* Candidate index for the sensors of one rules phase (oracle_rules builds it, uses it and drops it): the bounding box of every living
 * organism and the dead cells bucketed in a coarse grid. Nothing the index is built from changes during the phase (positions move in
 * Space.step, dead cells are created in settle; a cell that dies becomes DYING, not absent), every candidate still goes through the exact
 * predicate below, and eye / olfactory encodings are maxima over the view (order-free): results are bit-identical to the full scan, which
 * stays the path of every call outside a rules phase (think_now, the sensor-view probe) and of ORACLE_BRUTE=1. At 100k organisms the full
 * scan is 1e11 predicate evaluations per tick. */
#define VI_CELL 1024.0
typedef struct { const SapioWorld *w; double *bb; int gnx, gny; int *dstart; int *dlist; double rmax; } ViewIndex;
static __thread ViewIndex g_vi = {0};       /* per thread: independent worlds may be ticked from several threads of one process */
void oracle_view_index_drop(void) { free(g_vi.bb); free(g_vi.dstart); free(g_vi.dlist); memset(&g_vi, 0, sizeof(g_vi)); }
static int vi_cell(double v, int n) { int c = (int)floor(v / VI_CELL); return c < 0 ? 0 : (c >= n ? n - 1 : c); }
void oracle_view_index_build(const SapioWorld *w) {
    oracle_view_index_drop();
    const char *brute = getenv("ORACLE_BRUTE");
    if (brute && brute[0] == '1') return;
    const SapioParams *p = &w->p;
    g_vi.bb = (double *)malloc(sizeof(double) * 4 * (size_t)(p->n_org > 0 ? p->n_org : 1));
    for (int o = 0; o < p->n_org; o++) {
        double *b = g_vi.bb + 4 * (size_t)o;
        b[0] = b[1] = INFINITY; b[2] = b[3] = -INFINITY;
        if (w->org_state[o] != 1) continue;
        for (int k = 0; k < w->org_ncells[o]; k++) {
            int r2 = o * MAXC + k;
            if (w->c_alive[r2] == 0) continue;
            double x = w->c_pos[2 * r2], y = w->c_pos[2 * r2 + 1], rad = w->c_size[r2] * 0.5;
            if (x - rad < b[0]) b[0] = x - rad;
            if (y - rad < b[1]) b[1] = y - rad;
            if (x + rad > b[2]) b[2] = x + rad;
            if (y + rad > b[3]) b[3] = y + rad;
        }
    }
    double xmax = p->width, ymax = p->height;
    if (p->slab_mode && p->slab_x1 > xmax) xmax = p->slab_x1;
    g_vi.gnx = (int)ceil((xmax > 0 ? xmax : 1.0) / VI_CELL) + 1; g_vi.gny = (int)ceil((ymax > 0 ? ymax : 1.0) / VI_CELL) + 1;
    const size_t ng = (size_t)g_vi.gnx * g_vi.gny;
    g_vi.dstart = (int *)calloc(ng + 1, sizeof(int));
    g_vi.dlist = (int *)malloc(sizeof(int) * (size_t)(p->n_dead > 0 ? p->n_dead : 1));
    g_vi.rmax = 0;
    for (int d = 0; d < p->n_dead; d++) {
        if (!w->d_state[d]) continue;
        g_vi.dstart[(size_t)vi_cell(w->d_pos[2 * d + 1], g_vi.gny) * g_vi.gnx + vi_cell(w->d_pos[2 * d], g_vi.gnx) + 1]++;
        double rad = w->d_size[d] * 0.5; if (rad > g_vi.rmax) g_vi.rmax = rad;
    }
    for (size_t c = 0; c < ng; c++) g_vi.dstart[c + 1] += g_vi.dstart[c];
    int *fill = (int *)malloc(sizeof(int) * (ng + 1));
    memcpy(fill, g_vi.dstart, sizeof(int) * (ng + 1));
    for (int d = 0; d < p->n_dead; d++) {
        if (!w->d_state[d]) continue;
        g_vi.dlist[fill[(size_t)vi_cell(w->d_pos[2 * d + 1], g_vi.gny) * g_vi.gnx + vi_cell(w->d_pos[2 * d], g_vi.gnx)]++] = d;
    }
    free(fill);
    g_vi.w = w;
}

static void for_each_viewed(const SapioWorld *w, int o, int row, double R, view_cb cb, void *ctx) {
    const SapioParams *p = &w->p;
    double sx = w->c_spos[2 * row], sy = w->c_spos[2 * row + 1];
    const int indexed = g_vi.w == w && g_vi.bb != NULL;
    for (int oo = 0; oo < p->n_org; oo++) {
        if (oo == o || w->org_state[oo] != 1) continue;
        if (indexed) {                                                          /* the disc cannot reach the organism's box */
            const double *b = g_vi.bb + 4 * (size_t)oo;
            if (sx + R < b[0] || sx - R > b[2] || sy + R < b[1] || sy - R > b[3]) continue;
        }
        for (int k = 0; k < w->org_ncells[oo]; k++) {
            int r2 = oo * MAXC + k;
            if (w->c_alive[r2] == 0) continue;
            double x = w->c_pos[2 * r2], y = w->c_pos[2 * r2 + 1], rad = w->c_size[r2] * 0.5;
            double dx = x - sx, dy = y - sy, md = R + rad;
            if (dx * dx + dy * dy < md * md) cb(ctx, (uint32_t)r2, x, y, rad, 1);
        }
    }
    if (indexed) {
        const double reach = R + g_vi.rmax;
        const int cx0 = vi_cell(sx - reach, g_vi.gnx), cx1 = vi_cell(sx + reach, g_vi.gnx), cy0 = vi_cell(sy - reach, g_vi.gny), cy1 = vi_cell(sy + reach, g_vi.gny);
        for (int cy = cy0; cy <= cy1; cy++)
            for (int q = g_vi.dstart[(size_t)cy * g_vi.gnx + cx0], e = g_vi.dstart[(size_t)cy * g_vi.gnx + cx1 + 1]; q < e; q++) {
                const int d = g_vi.dlist[q];
                if (!w->d_state[d] || w->d_owner[d] == w->org_uid[o]) continue;
                double x = w->d_pos[2 * d], y = w->d_pos[2 * d + 1], rad = w->d_size[d] * 0.5;
                double dx = x - sx, dy = y - sy, md = R + rad;
                if (dx * dx + dy * dy < md * md) cb(ctx, (uint32_t)(p->n_org * MAXC + d), x, y, rad, 0);
            }
        return;
    }
    for (int d = 0; d < p->n_dead; d++) {
        if (!w->d_state[d] || w->d_owner[d] == w->org_uid[o]) continue;
        double x = w->d_pos[2 * d], y = w->d_pos[2 * d + 1], rad = w->d_size[d] * 0.5;
        double dx = x - sx, dy = y - sy, md = R + rad;
        if (dx * dx + dy * dy < md * md) cb(ctx, (uint32_t)(p->n_org * MAXC + d), x, y, rad, 0);
    }
}

typedef struct { double px, py, R, rot; double eye[8]; double smell; uint32_t *hits; int nhits, cap; } SenseCtx;

/* _adv_get_eye_input (neural.py:123-196, SURVEY E-5) */
static void eye_cb(void *vctx, uint32_t body, double x, double y, double r, int alive) {
    SenseCtx *c = (SenseCtx *)vctx; (void)alive;
    if (c->hits && c->nhits < c->cap) c->hits[c->nhits] = body;
    c->nhits++;
    double ang = atan2(y - c->py, x - c->px) - c->rot;
    double dx = cos(ang), dy = sin(ang);
    double dist = sqrt((x - c->px) * (x - c->px) + (y - c->py) * (y - c->py));
    double val = fabs((c->R - (dist + r)) / c->R);
    int idx;
    if (dx >= 0.0) { if (dy < 0.0) idx = dx < .707 ? 0 : 1; else idx = dx >= .707 ? 2 : 3; }
    else { if (dy >= 0.0) idx = dy >= .707 ? 4 : 5; else idx = dx < -.707 ? 6 : 7; }
    if (c->eye[idx] < val) c->eye[idx] = val;
}
/* _get_olfactory_input (neural.py:197-218) */
static void smell_cb(void *vctx, uint32_t body, double x, double y, double r, int alive) {
    SenseCtx *c = (SenseCtx *)vctx;
    if (c->hits && c->nhits < c->cap) c->hits[c->nhits] = body;
    c->nhits++;
    if (alive) return;
    double dist = sqrt((x - c->px) * (x - c->px) + (y - c->py) * (y - c->py));
    double val = fabs((c->R - (dist + r)) / c->R);
    if (val > c->smell) c->smell = val;
}

/* parity probe: the view (hit list) of sensor cell `row`; returns the member count */
int oracle_sensor_view(const SapioWorld *w, int o, int k, uint32_t *hits, int cap) {
    int row = o * MAXC + k, t = w->c_type[row];
    SenseCtx c; memset(&c, 0, sizeof(c));
    c.hits = hits; c.cap = cap; c.px = w->c_pos[2 * row]; c.py = w->c_pos[2 * row + 1];
    c.R = (double)w->c_size[row] * (t == SAPIO_T_EYE ? w->p.eyesight_mult : w->p.olfactory_mult);
    for_each_viewed(w, o, row, c.R, t == SAPIO_T_EYE ? eye_cb : smell_cb, &c);
    return c.nhits;
}

/* builds the (unrounded) input vector of organism o; returns its length */
int oracle_inputs(const SapioWorld *w, int o, const double *energy, const uint8_t *alive, double rotation_deg, double *in) {
    int n = 0;
    for (int q = 0; q < w->org_nincell[o]; q++) {
        int k = w->org_incell[o * MAXC + q], row = o * MAXC + k, t = w->c_type[row];
        SenseCtx c; memset(&c, 0, sizeof(c));
        c.px = w->c_pos[2 * row]; c.py = w->c_pos[2 * row + 1];
        c.rot = rotation_deg * (M_PI / 180.0);
        if (t == SAPIO_T_EYE) {
            c.R = (double)w->c_size[row] * w->p.eyesight_mult;                  /* sight_range (sprites.py:301) */
            if (alive[k] == 1) for_each_viewed(w, o, row, c.R, eye_cb, &c);
            for (int i = 0; i < 8; i++) in[n++] = c.eye[i];
        } else {
            c.R = (double)w->c_size[row] * w->p.olfactory_mult;                 /* smell_range (sprites.py:313) */
            if (alive[k] == 1) for_each_viewed(w, o, row, c.R, smell_cb, &c);
            in[n++] = c.smell;
        }
    }
    in[n++] = w->org_move[4 * o + 1];                                            /* speed, x_direction, y_direction (neural.py:239-244) */
    in[n++] = w->org_move[4 * o + 2];
    in[n++] = w->org_move[4 * o + 3];
    double cur = 0, mx = 0;                                                      /* _get_energy_input (neural.py:227-233) */
    for (int k = 0; k < w->org_ncells[o]; k++) { if (alive[k] == 1) cur += energy[k]; mx += w->c_storage[o * MAXC + k]; }
    in[n++] = (cur / mx * 100.0) / 100.;
    return n;
}

/* pinning hook: the five uniforms (boredom gate, 4 for randn) the reference was fed, instead of the Philox draws */
static const float *g_think_override = NULL;
void oracle_set_think_uniforms(const float *u5) { g_think_override = u5; }

/* neural.activate. History/STM are rings indexed by a running count (slot = count % capacity). */
int oracle_think(SapioWorld *w, int o, const double *health, const double *energy, const uint8_t *alive, double rotation_deg) {
    (void)health;
    const SapioParams *p = &w->p;
    const int IN = p->in_cap;
    const int32_t *ldim = w->org_ldim + o * 16;
    int nl = w->org_nlayers[o], I = ldim[0];
    double raw[1024];
    float x[1024], out[4];
    int n = oracle_inputs(w, o, energy, alive, rotation_deg, raw);
    if (n != I) return -1;
    for (int i = 0; i < I; i++) x[i] = rintf((float)raw[i] * 1000.0f) / 1000.0f;      /* neural.py:285 (torch.round: half-even) */
    const float *brain = w->org_brain + (size_t)o * p->brain_cap;
    const int H = w->org_rnn_h[o], RC = p->rnn_cap;
    float *lasth = w->org_last_hidden + (size_t)o * RC, hid[SAPIO_RNN_MAXH];
    for (int i = 0; i < H; i++) x[I + i] = lasth[i];                                   /* network.forward(flat_inputs, network.lastHidden), neural.py:287-288 */
    oracle_forward_rnn(brain, ldim, nl, H, x, out, NULL, hid);
    for (int i = 0; i < H; i++) lasth[i] = hid[i];
    for (int i = 0; i < 4; i++) w->org_out_raw[4 * o + i] = out[i];
    for (int i = 0; i < 4; i++) out[i] = rintf(out[i] * 1000.0f) / 1000.0f;            /* neural.py:292 */
    int32_t T = (int32_t)w->g_tick[0];
    uint32_t r[4];
    sapio_rng4(p->seed, w->org_uid[o], T, SAPIO_RNG_BOREDOM, 0, r);
    double gate_u = g_think_override ? (double)g_think_override[0] : (double)sapio_u01(r[0]);
    if (gate_u <= (double)w->org_boredom[o]) {                                        /* neural.py:294-295 */
        sapio_rng4(p->seed, w->org_uid[o], T, SAPIO_RNG_RANDN, 0, r);
        if (g_think_override) for (int q = 0; q < 4; q++) r[q] = (uint32_t)lrint((double)g_think_override[1 + q] * 16777216.0) << 8;
        for (int h = 0; h < 2; h++) {                                                  /* Box-Muller on (0,1] x [0,1) */
            double u1 = ((double)(r[2 * h] >> 8) + 1.0) / 16777216.0, u2 = (double)sapio_u01(r[2 * h + 1]);
            double rad = sqrt(-2.0 * log(u1)), th = 2.0 * M_PI * u2;
            out[2 * h] = (float)(rad * cos(th)); out[2 * h + 1] = (float)(rad * sin(th));
        }
    }
    /* previousInputs/Outputs/Dopamine append (neural.py:301-303); list before the append has nv entries */
    int cnt = w->org_hist_n[o], nv = cnt < SAPIO_HIST ? cnt : SAPIO_HIST;
    float *hin = w->org_hist_in + (size_t)o * SAPIO_HIST * IN, *hout = w->org_hist_out + o * 40, *hdo = w->org_hist_dopa + o * 10;
    double dopamine = w->org_dopamine[o], pain = w->org_pain[o];
    if ((fabs(dopamine) > 0.1 || fabs(pain) > 0.1) && nv + 1 > 3) {                    /* neural.py:305-311 */
        int s4 = (cnt - 3) % SAPIO_HIST;                                               /* previousInputs[-4] == old[nv-3] */
        int sc = w->org_stm_n[o], slot = sc % SAPIO_STM;
        float *sin_ = w->org_stm_in + ((size_t)o * SAPIO_STM + slot) * IN;
        memcpy(sin_, hin + (size_t)s4 * IN, sizeof(float) * (size_t)I);
        memcpy(w->org_stm_out + (o * SAPIO_STM + slot) * 4, hout + s4 * 4, sizeof(float) * 4);
        double d2 = hdo[(cnt - 2) % SAPIO_HIST], d1 = hdo[(cnt - 1) % SAPIO_HIST];
        w->org_stm_rew[o * SAPIO_STM + slot] = (float)((((0.0 + d2) + d1) + dopamine) / 3);
        w->org_stm_n[o] = sc + 1;
        if (H > 0) {                                                                   /* hidden_memory: append, and pop with the short-term memory (neural.py:312-318) */
            int32_t *hn = w->org_hm_n + 2 * o;
            memcpy(w->org_hm + ((size_t)o * SAPIO_HM + (size_t)(hn[0] % SAPIO_HM)) * RC, hid, sizeof(float) * (size_t)H);
            hn[0]++;
            if (sc + 1 > SAPIO_STM) hn[1]++;
        }
    }
    int slot = cnt % SAPIO_HIST;
    memcpy(hin + (size_t)slot * IN, x, sizeof(float) * (size_t)I);
    memcpy(hout + slot * 4, out, sizeof(float) * 4);
    hdo[slot] = (float)dopamine;
    w->org_hist_n[o] = cnt + 1;
    /* outputs -> movement, clamped to [-1, 1] (neural.py:320-349) */
    for (int i = 0; i < 4; i++) { float v = out[i]; v = v > 1.0f ? 1.0f : (v < -1.0f ? -1.0f : v); w->org_move[4 * o + i] = v; }
    /* calculateStimulation (neural.py:60-75): fp32 tensor arithmetic, oldest entry first */
    int nn = (cnt + 1) < SAPIO_HIST ? (cnt + 1) : SAPIO_HIST;
    float stim_sum = 0.0f;
    for (int i = 0; i < I; i++) {
        float s = 0.0f;
        for (int e = 0; e < nn; e++) { int sl = (cnt + 1 - nn + e) % SAPIO_HIST; s = s + hin[(size_t)sl * IN + i]; }
        float avg = s / (float)nn;
        stim_sum = stim_sum + fabsf(avg);
    }
    double new_stim = (double)(stim_sum / (float)I);
    float *sh = w->org_stim_hist + 10 * o;                                             /* neural.py:354-357 */
    for (int i = 0; i < 9; i++) sh[i] = sh[i + 1];
    sh[9] = (float)new_stim;
    double avg = 0; for (int i = 0; i < 10; i++) avg += (double)sh[i];
    avg /= 10;
    double stimulation = fabs(new_stim - avg);
    double cur = w->org_curiosity[o];
    w->org_stimulation[o] = (float)stimulation;
    w->org_boredom[o] = (float)((stimulation != 0 && cur != 0) ? (1.0 - stimulation) * cur : 0.0);
    return 0;
}

/* Trainer.train_rnn (networks.py:56-73) as neural.train_network calls it (neural.py:407-413), with everything it does restated:
 *  - the model runs over every memory row but only the LAST result is kept: input = the last row of trainingInput (list order: the
 *    most recently appended), hidden = trainingHidden[M - 1] -- the (M-1)-th entry EVER appended, since remove_memory never pops
 *    that list (neural.py:83-87);
 *  - targetData.resize_(result.shape) leaves the FIRST row of trainingOutput as the target;
 *  - MSELoss(mean) over the 4 outputs, backward, then plain SGD  p += -lr * grad  on model.parameters() that have a gradient: the
 *    input and output layers (the hidden layers sit in a Python list, networks.py:143; outputLayerH gets no gradient);
 *  - side effect: network.lastHidden = outputLayerH of that last row under the old weights. */
static void oracle_train_rnn(SapioWorld *w, int o, int M) {
    const SapioParams *p = &w->p;
    const int IN = p->in_cap, RC = p->rnn_cap, H = w->org_rnn_h[o];
    const int32_t *ldim = w->org_ldim + o * 16, *seq = w->org_mem_seq + (size_t)o * SAPIO_MEMROWS;
    const int nl = w->org_nlayers[o], I = ldim[0], F0 = I + H, h0 = ldim[1], hl = ldim[nl - 1];
    int last = 0, first = 0;
    for (int m = 1; m < M; m++) { if (seq[m] > seq[last]) last = m; if (seq[m] < seq[first]) first = m; }
    float x[1024], act[SAPIO_MAXL + 1][1024], out[4], hid[SAPIO_RNN_MAXH];
    memcpy(x, w->org_mem_in + ((size_t)o * SAPIO_MEMROWS + last) * IN, sizeof(float) * (size_t)I);
    for (int i = 0; i < H; i++) x[I + i] = (M - 1 < w->org_th_n[o]) ? w->org_th[((size_t)o * SAPIO_MEMROWS + (M - 1)) * RC + i] : 0.0f;
    float *brain = w->org_brain + (size_t)o * p->brain_cap;
    const float *cur = x; int off = 0;
    for (int l = 0; l < nl + 1; l++) {
        int fin = l == nl ? hl : ldim[l] + (l == 0 ? H : 0), fout = l == nl ? H : ldim[l + 1];
        const float *W = brain + off, *b = W + fout * fin;
        const float *src = l == nl ? act[nl - 1] : cur;
        float *dst = l == nl - 1 ? out : l == nl ? hid : act[l + 1];
        if (nl == 1 && l == nl) src = x;
        for (int r = 0; r < fout; r++) {
            float s = 0.0f;
            for (int c = 0; c < fin; c++) s += W[r * fin + c] * src[c];
            s += b[r];
            if (l == 0) s = 1.0f / (1.0f + expf(-s));
            dst[r] = s;
        }
        if (l < nl - 1) cur = dst;
        off += fout * fin + fout;
    }
    const float *y = w->org_mem_out + ((size_t)o * SAPIO_MEMROWS + first) * 4, *Hl = act[nl - 1], *H0 = act[1];
    float dO[4]; double loss = 0;
    for (int q = 0; q < 4; q++) { float e = out[q] - y[q]; loss += (double)e * e; dO[q] = 2.0f * e / 4.0f; }
    w->org_loss[o] = (float)(loss / 4);
    float *Win = brain, *bin = brain + h0 * F0, *Wout = brain + layer_off(ldim, nl - 1, H), *bout = Wout + 4 * hl;
    float d[2][1024]; int cd = 0;
    for (int c = 0; c < hl; c++) { float s = 0; for (int q = 0; q < 4; q++) s += dO[q] * Wout[q * hl + c]; d[0][c] = s; }
    for (int l = nl - 2; l >= 1; l--) {
        int fin = ldim[l], fout = ldim[l + 1];
        const float *W = brain + layer_off(ldim, l, H);
        for (int c = 0; c < fin; c++) { float s = 0; for (int r = 0; r < fout; r++) s += d[cd][r] * W[r * fin + c]; d[cd ^ 1][c] = s; }
        cd ^= 1;
    }
    const float lr = p->learning_rate;
    for (int q = 0; q < 4; q++) {
        for (int c = 0; c < hl; c++) { float g = dO[q] * Hl[c]; Wout[q * hl + c] = Wout[q * hl + c] + (-lr) * g; }
        bout[q] = bout[q] + (-lr) * dO[q];
    }
    for (int r = 0; r < h0; r++) {
        float dz = d[cd][r] * H0[r] * (1.0f - H0[r]);
        for (int c = 0; c < F0; c++) { float g = dz * x[c]; Win[r * F0 + c] = Win[r * F0 + c] + (-lr) * g; }
        bin[r] = bin[r] + (-lr) * dz;
    }
    for (int i = 0; i < H; i++) w->org_last_hidden[(size_t)o * RC + i] = hid[i];
}

/* neural.train_network (neural.py:369-423) for organism o.
 * Storage policy for the curated memory (a Python list in the reference): a replaced key is overwritten in place and the
 * trim moves the last row into the hole, instead of list.pop + append. Keys are unique, so only the fp summation order of the
 * batch and the tie-break of equal minimal rewards can tell the difference. */
int oracle_train_one(SapioWorld *w, int o) {
    const SapioParams *p = &w->p;
    const int IN = p->in_cap;
    const int32_t *ldim = w->org_ldim + o * 16;
    int nl = w->org_nlayers[o], I = ldim[0];
    int sc = w->org_stm_n[o], sn = sc < SAPIO_STM ? sc : SAPIO_STM;
    float *min_ = w->org_mem_in + (size_t)o * SAPIO_MEMROWS * IN, *mout = w->org_mem_out + (size_t)o * SAPIO_MEMROWS * 4;
    float *mrew = w->org_mem_rew + (size_t)o * SAPIO_MEMROWS;
    int M = w->org_mem_n[o];
    const int H = w->org_rnn_h[o], RC = p->rnn_cap;
    int32_t *seq = w->org_mem_seq + (size_t)o * SAPIO_MEMROWS;
    for (int i = 0; i < sn; i++) {                                                     /* chronological */
        int si = (sc - sn + i) % SAPIO_STM;
        const float *xin = w->org_stm_in + ((size_t)o * SAPIO_STM + si) * IN;
        /* best action among short-term memories with the same 1-dp key; sorted(...)[-1] keeps the LAST of equal maxima */
        int best = si;
        double best_r = -INFINITY;
        for (int j = 0; j < sn; j++) {
            int sj = (sc - sn + j) % SAPIO_STM;
            const float *xj = w->org_stm_in + ((size_t)o * SAPIO_STM + sj) * IN;
            int same = 1;
            for (int c = 0; c < I && same; c++) same = key_stm(xj[c]) == key_stm(xin[c]);
            if (!same) continue;
            double rj = w->org_stm_rew[o * SAPIO_STM + sj];
            if (rj >= best_r) { best_r = rj; best = sj; }
        }
        int add = 1, found = -1;
        for (int m = 0; m < M && found < 0; m++) {                                      /* rounded_training_data.index (first match) */
            int same = 1;
            for (int c = 0; c < I && same; c++) same = key_mem(min_[(size_t)m * IN + c]) == key_stm(xin[c]);
            if (same) found = m;
        }
        int dst = M;
        if (found >= 0) { if ((double)mrew[found] > best_r) add = 0; else dst = found; }
        if (add) {
            if (dst == M) { if (M >= SAPIO_MEMROWS) { w->g_stats[SAPIO_STAT_OVERFLOW]++; continue; } M++; }      /* the row table is full (only without finite_memory): dropped and counted */
            memcpy(min_ + (size_t)dst * IN, xin, sizeof(float) * (size_t)I);          /* roundList(mem[0], 3) == the fp32 value again */
            memcpy(mout + dst * 4, w->org_stm_out + (o * SAPIO_STM + best) * 4, sizeof(float) * 4);
            mrew[dst] = w->org_stm_rew[o * SAPIO_STM + si];                             /* sic: mem[2], not bestReward (neural.py:395) */
            seq[dst] = w->org_mem_stamp[o]++;                                           /* remove_memory + append: the row is now the LAST of the list */
            if (H > 0) {                                                                /* trainingHidden.append(hidden_memory[i]) (neural.py:396-397) */
                int tn = w->org_th_n[o];
                if (tn < SAPIO_MEMROWS) {
                    const int32_t *hn = w->org_hm_n + 2 * o;
                    memcpy(w->org_th + ((size_t)o * SAPIO_MEMROWS + tn) * RC, w->org_hm + ((size_t)o * SAPIO_HM + (size_t)((hn[1] + i) % SAPIO_HM)) * RC, sizeof(float) * (size_t)H);
                    w->org_th_n[o] = tn + 1;
                }
            }
        }
    }
    while (M > p->memory_limit && (w->org_flags[o] & SAPIO_F_FINITE_MEMORY)) {          /* neural.py:399-404 */
        int lo = 0;
        for (int m = 1; m < M; m++) if (mrew[m] < mrew[lo]) lo = m;
        M--;
        if (lo != M) {
            memcpy(min_ + (size_t)lo * IN, min_ + (size_t)M * IN, sizeof(float) * (size_t)I);
            memcpy(mout + lo * 4, mout + M * 4, sizeof(float) * 4);
            mrew[lo] = mrew[M]; seq[lo] = seq[M];
        }
    }
    w->org_mem_n[o] = M;
    if (M > 0 && H > 0) oracle_train_rnn(w, o, M);
    else if (M > 0) {
        /* Trainer.train_epoch (networks.py:75-87): full batch, MSELoss(mean), Adam on input+output layers only */
        float *brain = w->org_brain + (size_t)o * p->brain_cap;
        int h0 = ldim[1], hl = ldim[nl - 1];
        int off_out = layer_off(ldim, nl - 1, 0);
        float *Win = brain, *bin = brain + h0 * I, *Wout = brain + off_out, *bout = Wout + 4 * hl;
        int np_in = h0 * I + h0, np_out = 4 * hl + 4;
        double *g = (double *)calloc((size_t)(np_in + np_out), sizeof(double));
        float *H = (float *)malloc(sizeof(float) * 1024 * (size_t)(nl + 1));
        double loss = 0;
        for (int m = 0; m < M; m++) {
            const float *x = min_ + (size_t)m * IN, *y = mout + m * 4;
            /* forward keeping every layer's activation */
            const float *cur = x; int off = 0;
            for (int l = 0; l < nl; l++) {
                int fin = ldim[l], fout = ldim[l + 1];
                const float *W = brain + off, *b = W + fout * fin;
                float *dst = H + 1024 * (size_t)(l + 1);
                for (int r = 0; r < fout; r++) {
                    float s = 0.0f;
                    for (int c = 0; c < fin; c++) s += W[r * fin + c] * cur[c];
                    s += b[r];
                    if (l == 0) s = 1.0f / (1.0f + expf(-s));
                    dst[r] = s;
                }
                cur = dst; off += fout * fin + fout;
            }
            const float *O = H + 1024 * (size_t)nl;
            float dO[4];
            for (int q = 0; q < 4; q++) { float e = O[q] - y[q]; loss += (double)e * e; dO[q] = 2.0f * e / (float)(M * 4); }
            const float *Hl = (nl >= 2) ? H + 1024 * (size_t)(nl - 1) : x;
            for (int q = 0; q < 4; q++) {
                for (int c = 0; c < hl; c++) g[np_in + q * hl + c] += (double)(dO[q] * Hl[c]);
                g[np_in + 4 * hl + q] += (double)dO[q];
            }
            /* back through the activation-free hidden chain to the first layer's output */
            float d[2][1024];
            int cur_d = 0;
            for (int c = 0; c < hl; c++) { float s = 0; for (int q = 0; q < 4; q++) s += dO[q] * Wout[q * hl + c]; d[0][c] = s; }
            for (int l = nl - 2; l >= 1; l--) {
                int fin = ldim[l], fout = ldim[l + 1];
                const float *W = brain + layer_off(ldim, l, 0);
                for (int c = 0; c < fin; c++) { float s = 0; for (int r = 0; r < fout; r++) s += d[cur_d][r] * W[r * fin + c]; d[cur_d ^ 1][c] = s; }
                cur_d ^= 1;
            }
            const float *H0 = H + 1024;
            for (int r = 0; r < h0; r++) {
                float dz = d[cur_d][r] * H0[r] * (1.0f - H0[r]);
                for (int c = 0; c < I; c++) g[r * I + c] += (double)(dz * x[c]);
                g[h0 * I + r] += (double)dz;
            }
        }
        w->org_loss[o] = (float)(loss / (M * 4));
        /* torch.optim.Adam defaults: betas (0.9, 0.999), eps 1e-8, no weight decay */
        int t = ++w->org_adam_t[o];
        float lr = p->learning_rate, b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
        double bc1 = 1.0 - pow((double)b1, t), bc2 = 1.0 - pow((double)b2, t);
        float step_size = (float)((double)lr / bc1), bc2s = (float)sqrt(bc2);
        float *am = w->org_adam_m + (size_t)o * p->adam_cap, *av = w->org_adam_v + (size_t)o * p->adam_cap;
        for (int i = 0; i < np_in + np_out; i++) {
            float gi = (float)g[i];
            float *prm = i < h0 * I ? &Win[i] : (i < np_in ? &bin[i - h0 * I] : (i < np_in + 4 * hl ? &Wout[i - np_in] : &bout[i - np_in - 4 * hl]));
            am[i] = b1 * am[i] + (1.0f - b1) * gi;
            av[i] = b2 * av[i] + (1.0f - b2) * gi * gi;
            float denom = sqrtf(av[i]) / bc2s + eps;
            *prm = *prm - step_size * (am[i] / denom);
        }
        free(g); free(H);
    }
    w->org_train_tick[o] = (int32_t)w->g_tick[0];
    w->g_stats[SAPIO_STAT_TRAINS]++;
    return 0;
}

int oracle_train(SapioWorld *w) {
    for (int o = 0; o < w->p.n_org; o++) {
        if (w->org_state[o] != 1 || !(w->org_req[o] & SAPIO_REQ_TRAIN)) continue;
        oracle_train_one(w, o);
    }
    return 0;
}
