/* genome.c -- CPU oracle: sprites.DNA restated (randomize, mutate, geometry helpers, build order).
 * TEST INFRASTRUCTURE ONLY (see oracle.h). Every function cites the reference lines it follows.
 * Python dict iteration order is insertion order; here it is the array order of OGenome.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* sprites.base_cell_info (sprites.py:67-121), index SAPIO_T_* */
static const double BASE_HEALTH[SAPIO_NTYPES]  = {30, 30, 8, 24, 15, 24, 24, 24, 15};
static const double BASE_IDLE[SAPIO_NTYPES]    = {1.2, 2.5, 2, 3.5, 3.6, 3, 1.2, 2.5, 2.5};
static const double BASE_USE[SAPIO_NTYPES]     = {1.2, 4, 2, 5, 3.6, 6, 1.2, 5, 2.5};
static const double BASE_STORAGE[SAPIO_NTYPES] = {15, 30, 20, 30, 33, 30, 30, 30, 33};
static const double BASE_DAMAGE[SAPIO_NTYPES]  = {0, 2, 0, 0, 0, 0, 0, 0, 0};

double usrc_next(USrc *u) {
    if (u->arr) {
        if (u->pos >= u->n_arr) { u->overrun = 1; return 0.5; }
        return (double)u->arr[u->pos++];
    }
    return (double)sapio_stream_u(&u->st);
}

/* random.randrange(a, b) / random.choice(seq) / random.triangular on top of one uniform each
 * (CPython Lib/random.py: triangular) */
static int u_randrange(USrc *u, int a, int b) { return a + (int)floor(usrc_next(u) * (double)(b - a)); }
static double u_triangular(USrc *u, double low, double mode, double high) {
    double x = usrc_next(u);
    if (high == low) return low;
    double c = (mode - low) / (high - low);
    if (x > c) { x = 1.0 - x; c = 1.0 - c; double t = low; low = high; high = t; }
    return low + (high - low) * sqrt(x * c);
}

void og_init(OGenome *g) { memset(g, 0, sizeof(*g)); }
void og_free(OGenome *g) { free(g->hw); g->hw = NULL; g->hw_len = 0; }
void og_copy(OGenome *dst, const OGenome *src) {   /* DNA.copy (sprites.py:410-411) */
    float *old = dst->hw;
    *dst = *src;
    dst->hw = NULL;
    free(old);
    if (src->hw_len > 0) {
        dst->hw = (float *)malloc(sizeof(float) * (size_t)src->hw_len);
        memcpy(dst->hw, src->hw, sizeof(float) * (size_t)src->hw_len);
    }
}
int og_index_of(const OGenome *g, int id) {
    for (int j = 0; j < g->n; j++) if (g->id[j] == id) return j;
    return -1;
}

/* finish_cell_info (sprites.py:238-253): (k*size) + (k*mass), the additive module-level version */
void og_finish_cell_info(OGenome *g, int j) {
    int t = g->type[j];
    double s = g->size[j], m = g->mass[j];
    g->maxhealth[j] = BASE_HEALTH[t] * s + BASE_HEALTH[t] * m;
    g->use_idle[j]  = BASE_IDLE[t] * s + BASE_IDLE[t] * m;
    g->use_act[j]   = BASE_USE[t] * s + BASE_USE[t] * m;
    g->damage[j]    = BASE_DAMAGE[t] * s + BASE_DAMAGE[t] * m;
    g->storage[j]   = BASE_STORAGE[t] * s + BASE_STORAGE[t] * m;
}

/* DNA._new_cell_id (sprites.py:528-532) */
static int og_new_cell_id(const OGenome *g) {
    int cid = 1;
    while (og_index_of(g, cid) >= 0) cid++;
    return cid;
}

/* DNA._closest_cell_to_point (sprites.py:705-718): min over cells of centre distance minus radius */
static double og_closest_dist(const OGenome *g, double x, double y) {
    double best = INFINITY;
    for (int j = 0; j < g->n; j++) {
        double d = sqrt(og_sq(g->relx[j] - x) + og_sq(g->rely[j] - y)) - g->size[j] / 2.0;
        if (d < best) best = d;
    }
    return best;
}

/* DNA._cell_mirrorable (sprites.py:732-750) */
static int og_cell_mirrorable(const OGenome *g, double x, double y, double size) {
    double r = size / 2.0, nx = x, ny = y;
    if (g->mirror_x) nx = -nx;
    if (g->mirror_y) ny = -ny;
    double distance = sqrt(og_sq(nx - x) + og_sq(ny - y)) - size / 2.0;
    double distance2 = og_closest_dist(g, nx, ny);
    return distance > r * 0.8 && distance2 > r * 0.8;
}

/* DNA._viable_cell_position (sprites.py:752-764); distanceThreshold = .8 (sprites.py:389) */
static int og_viable_cell_position(const OGenome *g, double x, double y, double size) {
    double dist = og_closest_dist(g, x, y);
    double r = size / 2.0;
    if (g->mirror_x || g->mirror_y)
        if (!og_cell_mirrorable(g, x, y, size)) return 0;
    return dist > r * 0.8;
}

/* DNA.add_randomized_cell + _make_random_cell (sprites.py:797-825, 551-582). Returns the new id, 0 on failure.
 * Draw order: size, mass, elasticity, friction, [type], then per try: grow_from, bearing. */
static int og_add_randomized_cell(OGenome *g, const SapioParams *p, USrc *u, int first_cell) {
    int cid = og_new_cell_id(g);
    double size = (double)u_randrange(u, p->size_lo, p->size_hi);
    double mass = (double)u_randrange(u, p->mass_lo, p->mass_hi);
    double elast = usrc_next(u);
    double fric = usrc_next(u);
    int type = first_cell ? SAPIO_T_HEART : u_randrange(u, 0, 8);   /* random.choice(cell_types) */
    double rx = 0.0, ry = 0.0, gx = 0.0, gy = 0.0;
    int parent = 0;
    if (g->n > 0 && !first_cell) {
        int iterations = 0, ok = 0;
        while (!ok) {
            if (iterations > 30) return 0;                             /* maximum_creation_tries (sprites.py:392) */
            iterations++;
            int from = u_randrange(u, 0, g->n);                        /* random.choice(list(self.cells)) */
            int deg = u_randrange(u, 0, 360);                          /* random_direction (sprites.py:124-129) */
            double rad = deg * (M_PI / 180.0);                         /* math.radians */
            gx = cos(rad); gy = sin(rad);
            /* _new_relative_pos (sprites.py:720-730) */
            double reach = g->size[from] / 2.1 + size / 2.1;
            rx = g->relx[from] + gx * reach;
            ry = g->rely[from] + gy * reach;
            parent = g->id[from];
            ok = og_viable_cell_position(g, rx, ry, size);
        }
    }
    if (g->n >= MAXC) return -1;                                        /* tile cap (ours; the reference has none) */
    int j = g->n++;
    g->id[j] = cid; g->type[j] = type; g->first[j] = first_cell;
    g->mirror_id[j] = first_cell ? cid : 0;                             /* sprites.py:567-570 */
    g->parent_id[j] = parent;
    g->size[j] = size; g->mass[j] = mass; g->elast[j] = elast; g->fric[j] = fric;
    g->relx[j] = rx; g->rely[j] = ry; g->growx[j] = gx; g->growy[j] = gy;
    og_finish_cell_info(g, j);
    return cid;
}

/* DNA._apply_mirror (sprites.py:766-795) */
static int og_apply_mirror(OGenome *g, int cell_id) {
    if (g->n >= MAXC) return -1;
    int j = og_index_of(g, cell_id);
    int new_id = og_new_cell_id(g);
    int pj = og_index_of(g, g->parent_id[j]);
    int k = g->n++;
    g->id[k] = new_id; g->type[k] = g->type[j]; g->first[k] = g->first[j];
    g->size[k] = g->size[j]; g->mass[k] = g->mass[j]; g->elast[k] = g->elast[j]; g->fric[k] = g->fric[j];
    g->maxhealth[k] = g->maxhealth[j]; g->use_idle[k] = g->use_idle[j]; g->use_act[k] = g->use_act[j];
    g->storage[k] = g->storage[j]; g->damage[k] = g->damage[j];       /* cell_info.copy(): stale stats travel */
    g->mirror_id[j] = new_id; g->mirror_id[k] = cell_id;
    g->relx[k] = g->mirror_x ? -g->relx[j] : g->relx[j];
    g->rely[k] = g->mirror_y ? -g->rely[j] : g->rely[j];
    g->growx[k] = g->mirror_x ? -g->growx[j] : g->growx[j];
    g->growy[k] = g->mirror_y ? -g->growy[j] : g->growy[j];
    g->parent_id[k] = g->mirror_id[pj];                                 /* new_grows_from (sprites.py:785) */
    return new_id;
}

/* number of network inputs: neural.setup_network (neural.py:449-467): 8 per eye, 1 per olfactory, speed, x, y, energy */
int og_input_size(const OGenome *g) {
    int n = 4;
    for (int j = 0; j < g->n; j++) {
        if (g->type[j] == SAPIO_T_EYE) n += 8;
        else if (g->type[j] == SAPIO_T_OLFACTORY) n += 1;
    }
    return n;
}
/* layer dims [in, h0, ..., h_last, 4]; networks.Network.setupLayers + expand_list (networks.py:9-15, 216-235) */
static int og_layer_dims(const OGenome *g, int dims[16]) {
    int n = 0;
    dims[n++] = og_input_size(g);
    for (int i = 0; i < g->n_hidden; i++) dims[n++] = g->hidden[i];
    dims[n++] = 4;
    return n;   /* linear layers = n - 1 */
}
int og_brain_params(const OGenome *g) {
    int dims[16], n = og_layer_dims(g, dims), tot = 0;
    for (int l = 0; l + 1 < n; l++) tot += dims[l + 1] * dims[l] + dims[l + 1];
    return tot;
}
int og_fits_caps(const OGenome *g, const SapioParams *p) {
    if (g->n > MAXC) return 0;
    const int H = p->rnn_hidden;                  /* RNN brain (networks.py:111-166): H more input columns, a second head of H rows */
    if (og_input_size(g) + H > p->in_cap) return 0;
    if (g->n_hidden < 1 || g->n_hidden + 1 > SAPIO_MAXL) return 0;
    for (int i = 0; i < g->n_hidden; i++) if (g->hidden[i] > p->width_cap) return 0;
    if (og_brain_params(g) + H * g->hidden[0] + H * g->hidden[g->n_hidden - 1] + H > p->brain_cap) return 0;
    return 1;
}

/* nn.Linear.reset_parameters restated: kaiming_uniform_(a=sqrt 5) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for W, same bound
 * for b; one uniform per element, W row-major then b. */
static float init_draw(USrc *u, int fan_in) {
    double x = usrc_next(u);
    return (float)((2.0 * x - 1.0) * (1.0 / sqrt((double)fan_in)));
}

/* the `tempNet = neural.setup_network(self)` of _setup_brain_structure / mutate_brain_layers: builds every layer (consumes
 * init draws for all of them in construction order: input, hidden..., output) and keeps the hidden weights only. */
static void og_tempnet_hidden_weights(OGenome *g, USrc *u_init) {
    int dims[16], n = og_layer_dims(g, dims);
    int len = 0;
    for (int l = 1; l + 2 < n; l++) len += dims[l + 1] * dims[l];
    free(g->hw);
    g->hw = (float *)malloc(sizeof(float) * (size_t)(len > 0 ? len : 1));
    g->hw_len = len;
    int off = 0;
    for (int l = 0; l + 1 < n; l++) {
        int fin = dims[l], fout = dims[l + 1];
        int hidden = (l >= 1 && l + 2 < n);
        for (int e = 0; e < fout * fin; e++) {
            float v = init_draw(u_init, fin);
            if (hidden) g->hw[off++] = v;
        }
        for (int e = 0; e < fout; e++) (void)init_draw(u_init, fin);
    }
}

/* DNA._setup_brain_structure (sprites.py:827-870) */
static void og_setup_brain_structure(OGenome *g, const SapioParams *p, USrc *u, USrc *u_init) {
    int input_size = 6;                                   /* _base_brain_structure (sprites.py:362-364) */
    for (int j = 0; j < g->n; j++) {
        if (g->type[j] == SAPIO_T_EYE) input_size += 4;  /* neural.output_lengths["eye"] */
        else if (g->type[j] == SAPIO_T_CARNIV) input_size += 1;
    }
    int hs = u_randrange(u, p->hidden_lo, p->hidden_hi);
    if (hs > SAPIO_MAXL - 1) hs = SAPIO_MAXL - 1;
    double dec[16];
    for (int i = 0; i < hs; i++)                           /* numpy.linspace(input_size, 4, hs) */
        dec[i] = (hs == 1) ? (double)input_size : (i == hs - 1 ? 4.0 : input_size + i * ((4.0 - input_size) / (hs - 1)));
    double mn = dec[0], mx = dec[0];
    for (int i = 1; i < hs; i++) { if (dec[i] < mn) mn = dec[i]; if (dec[i] > mx) mx = dec[i]; }
    double min_range = mn - 1.0, max_range = mx * 2.0;
    g->n_hidden = hs;
    for (int i = 0; i < hs; i++) {
        int h = (int)u_triangular(u, min_range, dec[i], max_range);
        while (h == 0) h = (int)u_triangular(u, min_range, dec[i], max_range);
        g->hidden[i] = h;
    }
    og_tempnet_hidden_weights(g, u_init);
}

/* DNA.randomize (sprites.py:872-913) */
int og_randomize(OGenome *g, const SapioParams *p, USrc *u, USrc *u_init) {
    float *keep = g->hw; og_init(g); g->hw = keep;
    /* numpy.random.choice([0,1], p=[1-m, m]) == (u >= 1-m) */
    g->mirror_x = usrc_next(u) >= (double)(1.0f - p->mirror_px) ? 1 : 0;
    g->mirror_y = usrc_next(u) >= (double)(1.0f - p->mirror_py) ? 1 : 0;
    g->curiosity = usrc_next(u);
    int count = u_randrange(u, p->cell_lo, p->cell_hi);
    int first = 1;
    for (int i = 0; i < count; i++) {
        int added = 0;
        while (!added) { added = og_add_randomized_cell(g, p, u, first); if (added < 0) return -1; }
        first = 0;
    }
    if (g->mirror_x || g->mirror_y) {
        int n0 = g->n;
        for (int j = 0; j < n0; j++)
            if (!g->first[j]) if (og_apply_mirror(g, g->id[j]) < 0) return -1;
    }
    og_setup_brain_structure(g, p, u, u_init);
    g->generation = 0;
    return 0;
}

/* DNA._lower_cells (sprites.py:447-454): cells without children, growth_pattern order */
static int og_lower_count(const OGenome *g) {
    int n = 0;
    for (int j = 0; j < g->n; j++) {
        int has = 0;
        for (int k = 0; k < g->n; k++) if (g->parent_id[k] == g->id[j] && k != j) { has = 1; break; }
        if (!has) n++;
    }
    return n;
}

/* DNA.mutate_cell_count (sprites.py:915-946). The removal branch calls new_dna.remove_cell(rCell) and drops the
 * returned copy (sprites.py:932), so it only consumes draws: len(lower_cells) per iteration (randomizeList). */
static int og_mutate_cell_count(OGenome *g, const SapioParams *p, double severity, USrc *u) {
    int current = g->n;
    int cr = (int)(severity * 18.0);
    if (!cr) return 0;
    int added = u_randrange(u, -cr, cr);
    while (current + added < 2) added = u_randrange(u, -cr, cr);
    if (added < 0) {
        int lower = og_lower_count(g);
        for (int i = 0; i < -added; i++)
            for (int k = 0; k < lower; k++) (void)usrc_next(u);
    } else if (added > 0) {
        for (int i = 0; i < added; i++) {
            int id = 0;
            if (g->n + ((g->mirror_x || g->mirror_y) ? 2 : 1) > MAXC) return -1;   /* tile cap (ours) */
            while (!id) { id = og_add_randomized_cell(g, p, u, 0); if (id < 0) return -1; }
            if (g->mirror_x || g->mirror_y) if (og_apply_mirror(g, id) < 0) return -1;
        }
    }
    return 0;
}

/* DNA.mutate_cell_types (sprites.py:948-961): type strings change, the stored derived stats do NOT (no finish_cell_info) */
static void og_mutate_cell_types(OGenome *g, double severity, USrc *u) {
    for (int j = 0; j < g->n; j++) {
        double x = usrc_next(u);
        if (x <= severity && !g->first[j]) {
            int t = u_randrange(u, 0, 8);
            g->type[j] = t;
            if (g->mirror_x || g->mirror_y) {
                int m = g->mirror_id[j];
                if (m) { int mj = og_index_of(g, m); if (mj >= 0) g->type[mj] = t; }
            }
        }
    }
}

/* DNA.mutate_cell_masses (sprites.py:963-978) */
static void og_mutate_cell_masses(OGenome *g, double severity, USrc *u) {
    for (int j = 0; j < g->n; j++) {
        if (usrc_next(u) <= severity) {
            int mr = (int)(severity * 10.0);
            if (!mr) return;
            int add = u_randrange(u, -mr, mr);
            while (g->mass[j] + add <= 0) add = u_randrange(u, -mr, mr);
            g->mass[j] += add;
        }
    }
}

/* Python 3 round(): half to even */
static int py_round(double x) { return (int)nearbyint(x); }

/* DNA.mutate_brain_layers (sprites.py:980-1045) */
static void og_mutate_brain_layers(OGenome *g, const SapioParams *p, double severity, int weight_persistence, USrc *u, USrc *u_fill, USrc *u_garb) {
    int old_n = g->n_hidden, old_h[16];
    memcpy(old_h, g->hidden, sizeof(old_h));
    int omin = old_h[0], omax = old_h[0];
    for (int i = 1; i < old_n; i++) { if (old_h[i] < omin) omin = old_h[i]; if (old_h[i] > omax) omax = old_h[i]; }
    int lc = py_round(severity * 10.0);
    if (lc <= 0) return;                                   /* randrange(-0, 0) raises ValueError -> return */
    (void)u_randrange(u, -lc, lc);                         /* the probing randrange consumes a draw (sprites.py:986) */
    int change = (int)u_triangular(u, -lc, 0, lc);
    int new_n, new_h[16];
    if (change > 0) new_n = change;                        /* sic: a positive change SETS the length (sprites.py:993-999) */
    else { new_n = old_n + change; if (new_n < 1) new_n = 1; }
    if (new_n > SAPIO_MAXL - 1) new_n = SAPIO_MAXL - 1;
    for (int i = 0; i < new_n; i++)
        new_h[i] = (i < old_n) ? old_h[i] : u_randrange(u, omin, omax + 1);
    int sc = py_round(severity * 10.0);
    for (int i = 0; i < new_n; i++) {
        int amt = (int)u_triangular(u, -sc, 0, sc);
        new_h[i] -= amt;
        if (new_h[i] < 2) new_h[i] = 2;
    }
    /* old hidden weights, per old layer: rows = old_h[l+1], cols = old_h[l] */
    int old_off[16], old_len[16], n_old_layers = old_n - 1 > 0 ? old_n - 1 : 0, o = 0;
    for (int l = 0; l < n_old_layers; l++) { old_off[l] = o; old_len[l] = old_h[l + 1] * old_h[l]; o += old_len[l]; }
    float *old_hw = g->hw; g->hw = NULL;
    g->n_hidden = new_n;
    memcpy(g->hidden, new_h, sizeof(int) * 16);
    int len = 0;
    for (int l = 0; l + 1 < new_n; l++) len += new_h[l + 1] * new_h[l];
    g->hw = (float *)malloc(sizeof(float) * (size_t)(len > 0 ? len : 1));
    g->hw_len = len;
    int off = 0;
    (void)p;
    for (int l = 0; l + 1 < new_n; l++) {
        int rows = new_h[l + 1], cols = new_h[l];
        for (int r = 0; r < rows; r++) {
            /* oldSubLayer = old_layers[l][r]; resize_(cols) on that row view re-reads the old matrix storage from the row's
             * offset (sprites.py:1038-1039); IndexError (no such layer/row) -> torch.rand(cols) (sprites.py:1040-1041).
             * Elements past the end of the old storage are uninitialised memory in the reference (resize_ grows the shared
             * storage): defined here as U(0,1) from a stream of their own, so the torch.rand stream stays aligned. */
            int have_row = weight_persistence && l < n_old_layers && r < old_h[l + 1];
            for (int c = 0; c < cols; c++) {
                float v;
                if (have_row) {
                    int idx = r * old_h[l] + c;
                    v = (idx < old_len[l]) ? old_hw[old_off[l] + idx] : (float)usrc_next(u_garb);
                } else v = (float)usrc_next(u_fill);
                g->hw[off++] = v;
            }
        }
    }
    free(old_hw);
}

/* DNA.mutate_curiosity (sprites.py:1047-1060) */
static void og_mutate_curiosity(OGenome *g, double severity, USrc *u) {
    double a = usrc_next(u), b = usrc_next(u);
    double c = g->curiosity + (a - b) * severity;
    if (c > 1.0) c = 1.0; else if (c < 0.0) c = 0.0;
    g->curiosity = c;
}

/* DNA.mutate (sprites.py:1068-1085). Returns 0, or -1 when the child would exceed the tile (ours). */
int og_mutate(OGenome *child, const OGenome *parent, const SapioParams *p, double severity, double neural_severity,
              int weight_persistence, USrc *u, USrc *u_fill, USrc *u_garb) {
    og_copy(child, parent);
    if (og_mutate_cell_count(child, p, severity, u) < 0) return -1;
    if (usrc_next(u) <= severity) og_mutate_cell_types(child, severity, u);
    og_mutate_cell_masses(child, severity, u);
    og_mutate_brain_layers(child, p, neural_severity, weight_persistence, u, u_fill, u_garb);
    og_mutate_curiosity(child, neural_severity, u);
    return 0;
}

/* DNA.get_rect (sprites.py:473-495): min/max are updated with the centre, not the edge (quirk C-13) */
void og_get_rect(const OGenome *g, double r[4]) {
    int f = 0;
    for (int j = 0; j < g->n; j++) if (g->first[j]) { f = j; break; }
    double minx = g->relx[f], maxx = g->relx[f], miny = g->rely[f], maxy = g->rely[f];
    for (int j = 0; j < g->n; j++) {
        double rad = g->size[j] / 2.0;
        if (g->relx[j] - rad < minx) minx = g->relx[j];
        if (g->relx[j] + rad > maxx) maxx = g->relx[j];
        if (g->rely[j] - rad < miny) miny = g->rely[j];
        if (g->rely[j] + rad > maxy) maxy = g->rely[j];
    }
    r[0] = minx; r[1] = miny; r[2] = maxx; r[3] = maxy;
}
/* DNA.get_size (sprites.py:497-512) */
void og_get_size(const OGenome *g, double s[2]) {
    double r[4];
    og_get_rect(g, r);
    double minx = r[0] < 0 ? -r[0] : r[0], miny = r[1] < 0 ? -r[1] : r[1];
    s[0] = r[2] + minx; s[1] = r[3] + miny;
}

/* Organism.build_body creation order (sprites.py:1704-1714 with _create_cell :1138-1162):
 * first cell, its direct children, then every genome cell not yet created followed by its direct children. */
int og_instance_order(const OGenome *g, int order[MAXC]) {
    int created[MAXC] = {0}, n = 0, f = -1;
    for (int j = 0; j < g->n; j++) if (g->first[j]) { f = j; break; }
    if (f < 0) return 0;
#define CREATE(j_) do { int jj = (j_); if (!created[jj]) { int pj = og_index_of(g, g->parent_id[jj]); \
        if (g->first[jj] || (pj >= 0 && created[pj])) { created[jj] = 1; order[n++] = jj; } } } while (0)
    for (int pass = 0; pass <= g->n; pass++) {
        int j = (pass == 0) ? f : pass - 1;
        if (created[j] && pass > 0) continue;
        int was = created[j];
        CREATE(j);
        if (!was && created[j])
            for (int k = 0; k < g->n; k++) if (g->parent_id[k] == g->id[j] && k != j) CREATE(k);
    }
#undef CREATE
    return n;
}
