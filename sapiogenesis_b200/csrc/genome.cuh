// genome.cuh -- sprites.DNA on the device: randomize (spawn recipe), mutate (reproduction), geometry helpers, build order.
// One warp per genome: the dict-ordered, draw-by-draw logic of the reference runs on lane 0 over a genome held in shared
// memory (births are rare events; the wide work -- 360-bearing placement test, brain initialisation, row writes -- is done by
// the whole warp in repro.cuh). Every draw is a Philox uniform (rng.cuh), consumed in the reference's order, so decisions
// are bit-exact against the oracle, which in turn is pinned against sprites.DNA.randomize / DNA.mutate on identical uniforms.
// Arithmetic is double like the reference's Python floats.
#pragma once
#include "common.cuh"
#include "rng.cuh"

struct DGenome {                      // == sprites.DNA (sprites.py:329-408), dict order == array order
    int n;
    int16_t id[MAXC], mirror_id[MAXC], parent_id[MAXC];        // dna ids (1-based); 0 = None
    int8_t type[MAXC], first[MAXC];
    double size[MAXC], mass[MAXC], elast[MAXC], fric[MAXC], relx[MAXC], rely[MAXC], growx[MAXC], growy[MAXC];
    double maxhealth[MAXC], use_idle[MAXC], use_act[MAXC], storage[MAXC], damage[MAXC];
    int mirror_x, mirror_y, n_hidden, generation;
    int hidden[16];
    double curiosity;
};

// sprites.base_cell_info (sprites.py:67-121), index SAPIO_T_*
__constant__ double C_BASE_HEALTH[SAPIO_NTYPES]  = {30, 30, 8, 24, 15, 24, 24, 24, 15};
__constant__ double C_BASE_IDLE[SAPIO_NTYPES]    = {1.2, 2.5, 2, 3.5, 3.6, 3, 1.2, 2.5, 2.5};
__constant__ double C_BASE_USE[SAPIO_NTYPES]     = {1.2, 4, 2, 5, 3.6, 6, 1.2, 5, 2.5};
__constant__ double C_BASE_STORAGE[SAPIO_NTYPES] = {15, 30, 20, 30, 33, 30, 30, 30, 33};
__constant__ double C_BASE_DAMAGE[SAPIO_NTYPES]  = {0, 2, 0, 0, 0, 0, 0, 0, 0};

__device__ __forceinline__ int g_randrange(RngStream &u, int a, int b) { return a + (int)floor((double)u.next() * (double)(b - a)); }
// random.triangular (CPython Lib/random.py)
__device__ __forceinline__ double g_triangular(RngStream &u, double low, double mode, double high) {
    double x = (double)u.next();
    if (high == low) return low;
    double c = (mode - low) / (high - low);
    if (x > c) { x = 1.0 - x; c = 1.0 - c; double t = low; low = high; high = t; }
    return low + (high - low) * sqrt(x * c);
}
__device__ __forceinline__ int g_index_of(const DGenome &g, int id) { for (int j = 0; j < g.n; j++) if (g.id[j] == id) return j; return -1; }
// DNA new cell id (sprites.py:528-532): the smallest positive id not in use. One pass over a bitmap instead of a search per candidate.
__device__ __forceinline__ int g_new_cell_id(const DGenome &g) {
    uint64_t used[4] = {0, 0, 0, 0};
    for (int j = 0; j < g.n; j++) { const int id = g.id[j]; if (id >= 0 && id < 256) used[id >> 6] |= 1ull << (id & 63); }
    used[0] |= 1ull;                                             // ids start at 1
    for (int q = 0; q < 4; q++) if (~used[q]) return q * 64 + __ffsll((long long)~used[q]) - 1;
    int cid = 256; while (g_index_of(g, cid) >= 0) cid++; return cid;
}
// finish_cell_info (sprites.py:238-253)
__device__ __forceinline__ void g_finish_cell_info(DGenome &g, int j) {
    int t = g.type[j]; double s = g.size[j], m = g.mass[j];
    g.maxhealth[j] = C_BASE_HEALTH[t] * s + C_BASE_HEALTH[t] * m;
    g.use_idle[j] = C_BASE_IDLE[t] * s + C_BASE_IDLE[t] * m;
    g.use_act[j] = C_BASE_USE[t] * s + C_BASE_USE[t] * m;
    g.damage[j] = C_BASE_DAMAGE[t] * s + C_BASE_DAMAGE[t] * m;
    g.storage[j] = C_BASE_STORAGE[t] * s + C_BASE_STORAGE[t] * m;
}
// DNA._closest_cell_to_point (sprites.py:705-718)
// The genome logic is scalar (one lane walks the reference's dict code), but its one hot loop -- the distance of a candidate
// position to every cell, in double, tried up to 31 times per added cell -- is served by the other lanes of the warp:
// lane 0 posts (x, y) in shared memory, everybody takes the cells lane, lane + 32, the partial minima come back through shared
// memory. min is exact in any order: the result is the serial loop's. Barriers are __syncwarp (legal from different code paths).
struct GenomeSvc { int cmd; double x, y; double part[32]; };      // cmd: 1 = closest distance, 0 = done
__device__ __forceinline__ double g_closest_part(const DGenome &g, double x, double y, int lane) {
    double best = INFINITY;
    for (int j = lane; j < g.n; j += 32) {
        double dx = g.relx[j] - x, dy = g.rely[j] - y;
        double d = sqrt(dx * dx + dy * dy) - g.size[j] / 2.0;
        if (d < best) best = d;
    }
    return best;
}
__device__ __forceinline__ void g_svc_worker(const DGenome &g, GenomeSvc *svc, int lane) {       // lanes 1..31
    for (;;) {
        __syncwarp();                                             // a request is posted
        if (svc->cmd == 0) break;
        svc->part[lane] = g_closest_part(g, svc->x, svc->y, lane);
        __syncwarp();                                             // partial minima are in
    }
}
__device__ __forceinline__ void g_svc_done(GenomeSvc *svc) { if (svc) { svc->cmd = 0; __syncwarp(); } }
__device__ __forceinline__ double g_closest_dist(const DGenome &g, double x, double y, GenomeSvc *svc = nullptr) {
    if (svc) {
        svc->cmd = 1; svc->x = x; svc->y = y;
        __syncwarp();
        double best = g_closest_part(g, x, y, 0);
        __syncwarp();
        for (int l = 1; l < 32; l++) { double d = svc->part[l]; if (d < best) best = d; }
        return best;
    }
    double best = INFINITY;
    for (int j = 0; j < g.n; j++) {
        double dx = g.relx[j] - x, dy = g.rely[j] - y;
        double d = sqrt(dx * dx + dy * dy) - g.size[j] / 2.0;
        if (d < best) best = d;
    }
    return best;
}
// DNA._viable_cell_position + _cell_mirrorable (sprites.py:732-764), distanceThreshold .8
__device__ __forceinline__ bool g_viable_cell_position(const DGenome &g, double x, double y, double size, GenomeSvc *svc = nullptr) {
    double dist = g_closest_dist(g, x, y, svc), r = size / 2.0;
    if (g.mirror_x || g.mirror_y) {
        double nx = g.mirror_x ? -x : x, ny = g.mirror_y ? -y : y;
        double ddx = nx - x, ddy = ny - y;
        double distance = sqrt(ddx * ddx + ddy * ddy) - size / 2.0;
        double distance2 = g_closest_dist(g, nx, ny, svc);
        if (!(distance > r * 0.8 && distance2 > r * 0.8)) return false;
    }
    return dist > r * 0.8;
}
// DNA.add_randomized_cell + _make_random_cell (sprites.py:797-825, 551-582): new id, 0 = no room after 31 tries, -1 = tile full
__device__ __forceinline__ int g_add_randomized_cell(DGenome &g, const SapioParams &p, RngStream &u, bool first_cell, GenomeSvc *svc = nullptr) {
    int cid = g_new_cell_id(g);
    double size = (double)g_randrange(u, p.size_lo, p.size_hi), mass = (double)g_randrange(u, p.mass_lo, p.mass_hi);
    double elast = (double)u.next(), fric = (double)u.next();
    int type = first_cell ? SAPIO_T_HEART : g_randrange(u, 0, 8);
    double rx = 0, ry = 0, gx = 0, gy = 0; int parent = 0;
    if (g.n > 0 && !first_cell) {
        int iterations = 0; bool ok = false;
        while (!ok) {
            if (iterations > 30) return 0;
            iterations++;
            int from = g_randrange(u, 0, g.n), deg = g_randrange(u, 0, 360);
            double rad = deg * (3.14159265358979323846 / 180.0);
            gx = cos(rad); gy = sin(rad);
            double reach = g.size[from] / 2.1 + size / 2.1;
            rx = g.relx[from] + gx * reach; ry = g.rely[from] + gy * reach;
            parent = g.id[from];
            ok = g_viable_cell_position(g, rx, ry, size, svc);
        }
    }
    if (g.n >= MAXC) return -1;
    int j = g.n++;
    g.id[j] = (int16_t)cid; g.type[j] = (int8_t)type; g.first[j] = first_cell; g.mirror_id[j] = (int16_t)(first_cell ? cid : 0); g.parent_id[j] = (int16_t)parent;
    g.size[j] = size; g.mass[j] = mass; g.elast[j] = elast; g.fric[j] = fric; g.relx[j] = rx; g.rely[j] = ry; g.growx[j] = gx; g.growy[j] = gy;
    g_finish_cell_info(g, j);
    return cid;
}
// DNA._apply_mirror (sprites.py:766-795)
__device__ __forceinline__ int g_apply_mirror(DGenome &g, int cell_id) {
    if (g.n >= MAXC) return -1;
    int j = g_index_of(g, cell_id), new_id = g_new_cell_id(g), pj = g_index_of(g, g.parent_id[j]), k = g.n++;
    g.id[k] = (int16_t)new_id; g.type[k] = g.type[j]; g.first[k] = g.first[j];
    g.size[k] = g.size[j]; g.mass[k] = g.mass[j]; g.elast[k] = g.elast[j]; g.fric[k] = g.fric[j];
    g.maxhealth[k] = g.maxhealth[j]; g.use_idle[k] = g.use_idle[j]; g.use_act[k] = g.use_act[j]; g.storage[k] = g.storage[j]; g.damage[k] = g.damage[j];
    g.mirror_id[j] = (int16_t)new_id; g.mirror_id[k] = (int16_t)cell_id;
    g.relx[k] = g.mirror_x ? -g.relx[j] : g.relx[j]; g.rely[k] = g.mirror_y ? -g.rely[j] : g.rely[j];
    g.growx[k] = g.mirror_x ? -g.growx[j] : g.growx[j]; g.growy[k] = g.mirror_y ? -g.growy[j] : g.growy[j];
    g.parent_id[k] = g.mirror_id[pj];
    return new_id;
}
// neural.setup_network input count (neural.py:449-467)
__device__ __forceinline__ int g_input_size(const DGenome &g) { int n = 4; for (int j = 0; j < g.n; j++) { if (g.type[j] == SAPIO_T_EYE) n += 8; else if (g.type[j] == SAPIO_T_OLFACTORY) n += 1; } return n; }
__device__ __forceinline__ int g_brain_params(const DGenome &g) {
    int prev = g_input_size(g), tot = 0;
    for (int i = 0; i < g.n_hidden; i++) { tot += g.hidden[i] * prev + g.hidden[i]; prev = g.hidden[i]; }
    return tot + 4 * prev + 4;
}
__device__ __forceinline__ bool g_fits_caps(const DGenome &g, const SapioParams &p) {
    const int H = p.rnn_hidden;                   // RNN brain (networks.py:111-166): H more input columns, a second head of H rows
    if (g.n > MAXC || g_input_size(g) + H > p.in_cap || g.n_hidden < 1 || g.n_hidden + 1 > SAPIO_MAXL) return false;
    for (int i = 0; i < g.n_hidden; i++) if (g.hidden[i] > p.width_cap) return false;
    return g_brain_params(g) + H * g.hidden[0] + H * g.hidden[g.n_hidden - 1] + H <= p.brain_cap;
}
// DNA._setup_brain_structure (sprites.py:827-866), structure only (the tempNet's weights are drawn by the caller)
__device__ __forceinline__ void g_setup_brain_structure(DGenome &g, const SapioParams &p, RngStream &u) {
    int input_size = 6;
    for (int j = 0; j < g.n; j++) { if (g.type[j] == SAPIO_T_EYE) input_size += 4; else if (g.type[j] == SAPIO_T_CARNIV) input_size += 1; }
    int hs = g_randrange(u, p.hidden_lo, p.hidden_hi);
    if (hs > SAPIO_MAXL - 1) hs = SAPIO_MAXL - 1;
    double dec[16];
    for (int i = 0; i < hs; i++) dec[i] = (hs == 1) ? (double)input_size : (i == hs - 1 ? 4.0 : input_size + i * ((4.0 - input_size) / (hs - 1)));
    double mn = dec[0], mx = dec[0];
    for (int i = 1; i < hs; i++) { if (dec[i] < mn) mn = dec[i]; if (dec[i] > mx) mx = dec[i]; }
    g.n_hidden = hs;
    for (int i = 0; i < hs; i++) {
        int h = (int)g_triangular(u, mn - 1.0, dec[i], mx * 2.0);
        while (h == 0) h = (int)g_triangular(u, mn - 1.0, dec[i], mx * 2.0);
        g.hidden[i] = h;
    }
}
// DNA.randomize (sprites.py:872-913); 0 ok, -1 tile overflow
__device__ __forceinline__ int g_randomize(DGenome &g, const SapioParams &p, RngStream &u) {
    g.n = 0;
    g.mirror_x = (double)u.next() >= (double)(1.0f - p.mirror_px) ? 1 : 0;
    g.mirror_y = (double)u.next() >= (double)(1.0f - p.mirror_py) ? 1 : 0;
    g.curiosity = (double)u.next();
    int count = g_randrange(u, p.cell_lo, p.cell_hi);
    bool first = true;
    for (int i = 0; i < count; i++) {
        int added = 0;
        while (!added) { added = g_add_randomized_cell(g, p, u, first); if (added < 0) return -1; }
        first = false;
    }
    if (g.mirror_x || g.mirror_y) { int n0 = g.n; for (int j = 0; j < n0; j++) if (!g.first[j]) if (g_apply_mirror(g, g.id[j]) < 0) return -1; }
    g_setup_brain_structure(g, p, u);
    g.generation = 0;
    return 0;
}
// DNA.mutate_* (sprites.py:915-1060) except the weight carry-over, which needs the parent's tensors (repro.cuh)
__device__ __forceinline__ int g_mutate_cells(DGenome &g, const SapioParams &p, double severity, RngStream &u, GenomeSvc *svc = nullptr) {
    {   // mutate_cell_count: the removal branch drops remove_cell()'s returned copy (sprites.py:932) -> draws only
        int current = g.n, cr = (int)(severity * 18.0);
        if (cr) {
            int added = g_randrange(u, -cr, cr);
            while (current + added < 2) added = g_randrange(u, -cr, cr);
            if (added < 0) {
                int lower = 0;
                for (int j = 0; j < g.n; j++) { bool has = false; for (int k = 0; k < g.n; k++) if (k != j && g.parent_id[k] == g.id[j]) { has = true; break; } if (!has) lower++; }
                for (int i = 0; i < -added; i++) for (int k = 0; k < lower; k++) (void)u.next();
            } else for (int i = 0; i < added; i++) {
                if (g.n + ((g.mirror_x || g.mirror_y) ? 2 : 1) > MAXC) return -1;
                int id = 0;
                while (!id) { id = g_add_randomized_cell(g, p, u, false, svc); if (id < 0) return -1; }
                if (g.mirror_x || g.mirror_y) if (g_apply_mirror(g, id) < 0) return -1;
            }
        }
    }
    if ((double)u.next() <= severity)   // mutate_cell_types: stored derived stats stay stale (no finish_cell_info)
        for (int j = 0; j < g.n; j++) {
            double x = (double)u.next();
            if (x <= severity && !g.first[j]) {
                int t = g_randrange(u, 0, 8);
                g.type[j] = (int8_t)t;
                if ((g.mirror_x || g.mirror_y) && g.mirror_id[j]) { int mj = g_index_of(g, g.mirror_id[j]); if (mj >= 0) g.type[mj] = (int8_t)t; }
            }
        }
    for (int j = 0; j < g.n; j++)       // mutate_cell_masses
        if ((double)u.next() <= severity) {
            int mr = (int)(severity * 10.0);
            if (!mr) break;
            int add = g_randrange(u, -mr, mr);
            while (g.mass[j] + add <= 0) add = g_randrange(u, -mr, mr);
            g.mass[j] += add;
        }
    return 0;
}
// mutate_brain_layers structure part (sprites.py:980-1020); old_h/old_n keep the parent's widths for the weight carry-over
__device__ __forceinline__ void g_mutate_brain_structure(DGenome &g, double severity, RngStream &u, int *old_h, int &old_n, bool &changed) {
    old_n = g.n_hidden; for (int i = 0; i < 16; i++) old_h[i] = g.hidden[i];
    changed = false;
    int lc = (int)nearbyint(severity * 10.0);
    if (lc <= 0) return;
    changed = true;
    int omin = old_h[0], omax = old_h[0];
    for (int i = 1; i < old_n; i++) { omin = min(omin, old_h[i]); omax = max(omax, old_h[i]); }
    (void)g_randrange(u, -lc, lc);
    int change = (int)g_triangular(u, -lc, 0, lc), new_n;
    if (change > 0) new_n = change; else { new_n = old_n + change; if (new_n < 1) new_n = 1; }
    if (new_n > SAPIO_MAXL - 1) new_n = SAPIO_MAXL - 1;
    int new_h[16];
    for (int i = 0; i < new_n; i++) new_h[i] = i < old_n ? old_h[i] : g_randrange(u, omin, omax + 1);
    for (int i = 0; i < new_n; i++) { int amt = (int)g_triangular(u, -lc, 0, lc); new_h[i] -= amt; if (new_h[i] < 2) new_h[i] = 2; }
    g.n_hidden = new_n; for (int i = 0; i < new_n; i++) g.hidden[i] = new_h[i];
}
__device__ __forceinline__ void g_mutate_curiosity(DGenome &g, double severity, RngStream &u) {
    double a = (double)u.next(), b = (double)u.next(), c = g.curiosity + (a - b) * severity;
    g.curiosity = c > 1.0 ? 1.0 : (c < 0.0 ? 0.0 : c);
}
// DNA.get_rect (sprites.py:473-495, quirk C-13) and get_size (sprites.py:497-512)
__device__ __forceinline__ void g_get_rect(const DGenome &g, double r[4]) {
    int f = 0; for (int j = 0; j < g.n; j++) if (g.first[j]) { f = j; break; }
    double minx = g.relx[f], maxx = minx, miny = g.rely[f], maxy = miny;
    for (int j = 0; j < g.n; j++) {
        double rad = g.size[j] / 2.0;
        if (g.relx[j] - rad < minx) minx = g.relx[j];
        if (g.relx[j] + rad > maxx) maxx = g.relx[j];
        if (g.rely[j] - rad < miny) miny = g.rely[j];
        if (g.rely[j] + rad > maxy) maxy = g.rely[j];
    }
    r[0] = minx; r[1] = miny; r[2] = maxx; r[3] = maxy;
}
__device__ __forceinline__ void g_get_size(const DGenome &g, double s[2]) {
    double r[4]; g_get_rect(g, r);
    s[0] = r[2] + (r[0] < 0 ? -r[0] : r[0]); s[1] = r[3] + (r[1] < 0 ? -r[1] : r[1]);
}
// Organism.build_body creation order (sprites.py:1704-1714, _create_cell :1138-1162)
__device__ __forceinline__ int g_instance_order(const DGenome &g, uint8_t *order) {
    uint64_t created = 0; int n = 0, f = -1;
    for (int j = 0; j < g.n; j++) if (g.first[j]) { f = j; break; }
    if (f < 0) return 0;
    for (int pass = 0; pass <= g.n; pass++) {
        int j = pass == 0 ? f : pass - 1;
        if ((created >> j) & 1ull) continue;
        int pj = g.first[j] ? -1 : g_index_of(g, g.parent_id[j]);
        if (!(g.first[j] || (pj >= 0 && ((created >> pj) & 1ull)))) continue;
        created |= 1ull << j; order[n++] = (uint8_t)j;
        for (int k = 0; k < g.n; k++) if (k != j && g.parent_id[k] == g.id[j] && !((created >> k) & 1ull)) { created |= 1ull << k; order[n++] = (uint8_t)k; }
    }
    return n;
}
// rebuild the DNA of a living organism from its rows (dna.cells order via c_gorder); lane 0 only
// whole warp: lane l loads the rows l and l + 32 (a single lane walking ~20 rows x 22 arrays of cold global memory, each row behind the load
// of its genome index, was most of a birth's latency); g in shared memory, __syncwarp() before anybody reads it
__device__ __forceinline__ void g_from_world(const SapioWorld &w, int org, DGenome &g) {
    const int nc = w.org_ncells[org], lane = lane_id();
    if (lane == 0) g.n = nc;
    for (int k = lane; k < nc; k += 32) {
        int row = org * MAXC + k, j = w.c_gorder[row];
        g.id[j] = w.c_id[row]; g.type[j] = (int8_t)w.c_type[row]; g.first[j] = w.c_parent[row] < 0;
        g.parent_id[j] = w.c_parent[row] < 0 ? 0 : w.c_id[org * MAXC + w.c_parent[row]];
        g.mirror_id[j] = w.c_mirror[row] < 0 ? 0 : w.c_id[org * MAXC + w.c_mirror[row]];
        g.size[j] = w.c_size[row]; g.mass[j] = w.c_mass[row]; g.elast[j] = w.c_elast[row]; g.fric[j] = w.c_fric[row];
        g.relx[j] = w.c_rel[2 * row]; g.rely[j] = w.c_rel[2 * row + 1]; g.growx[j] = w.c_grow[2 * row]; g.growy[j] = w.c_grow[2 * row + 1];
        g.maxhealth[j] = w.c_maxhealth[row]; g.use_idle[j] = w.c_use_idle[row]; g.use_act[j] = w.c_use_act[row]; g.storage[j] = w.c_storage[row]; g.damage[j] = w.c_damage[row];
    }
    if (lane == 0) {
        uint8_t fl = w.org_flags[org];
        g.mirror_x = (fl & SAPIO_F_MIRROR_X) != 0; g.mirror_y = (fl & SAPIO_F_MIRROR_Y) != 0;
        g.curiosity = w.org_curiosity[org]; g.generation = w.org_generation[org];
        g.n_hidden = w.org_nlayers[org] - 1;
    }
    const int nh = w.org_nlayers[org] - 1;
    for (int i = lane; i < nh; i += 32) g.hidden[i] = w.org_ldim[org * 16 + i + 1];
    __syncwarp();
}
