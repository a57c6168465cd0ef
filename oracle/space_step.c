/* space_step.c -- CPU oracle: pymunk.Space.step(dt) (physics.py:432) restated for the subset the reference uses, plus
 * Sprite.updateSprite friction (physics.py:250-275).
 * TEST INFRASTRUCTURE ONLY (see oracle.h). PARITY UNPINNED at the Chipmunk boundary: pymunk>=6.2.1 / Chipmunk2D 7.0.3 are
 * not in the reference checkout nor in this image; this follows the published algorithm of cpSpaceStep.c (step order),
 * cpBody.c (cpBodyUpdatePosition/Velocity), cpCollision.c (CircleToCircle, CircleToSegment), cpArbiter.c (PreStep,
 * ApplyCachedImpulse, ApplyImpulse) and cpPinJoint.c, as summarised in SURVEY.md Appendix A.
 *
 * Canonical solve order (Chipmunk's arbiter order depends on its BBTree/hash internals and is not reproducible; SURVEY A-7):
 *   arbiters   : (1) cross contacts -- pairs whose bodies are not two living cells of one organism, and not a living cell
 *                    against a wall -- sorted by (body_a, body_b);
 *                (2) per organism in slot order, its intra contacts (cell-cell, then cell-wall) sorted by (i, j);
 *   constraints: per organism in slot order (== creation order), sensor pins in row order, then pins (i,j) i<j
 *                lexicographic == insertion order of build_body (sprites.py:1716-1726).
 * Contacts are circle/circle or circle/segment, one point each, hash 0 => a pair touching on consecutive steps keeps its
 * jnAcc/jtAcc. A pair that separates restarts from zero (Chipmunk keeps the stale value for up to collisionPersistence=3
 * steps without applying it; that quirk is not reproduced).
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double x, y; } V2;
static inline V2 v2(double x, double y) { V2 r = {x, y}; return r; }
static inline double vdot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
static inline double vcross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
static inline V2 vperp(V2 a) { return v2(-a.y, a.x); }
static inline V2 vadd(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
static inline V2 vsub(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
static inline V2 vmul(V2 a, double s) { return v2(a.x * s, a.y * s); }
static inline V2 vrotate(V2 a, V2 b) { return v2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

typedef struct {
    V2 p, v, vb;
    double w, wb, m_inv, i_inv, r, e, u;
    int dynamic;
} Body;

typedef struct {
    int a, b;              /* indices into the body table */
    V2 n, r1, r2;
    double nMass, tMass, bias, bounce, e, u;
    double jn, jt, jBias;
} Contact;

typedef struct {
    int a, b;
    V2 n;
    double nMass, bias, jn;
} Pin;

/* SURVEY A-3: strict overlap test, evaluated in double on the fp32 positions (exact products) */
static int circles_overlap(double x1, double y1, double r1, double x2, double y2, double r2, double *dsq) {
    double dx = x2 - x1, dy = y2 - y1;
    double d2 = dx * dx + dy * dy, md = r1 + r2;
    *dsq = d2;
    return d2 < md * md;
}

/* cpCollision.c CircleToSegment for wall `wi` (physics.py:328-338: left, bottom, right, top; radius frame_width) */
static void wall_segment(const SapioParams *p, int wi, V2 *a, V2 *b) {
    double fw = p->frame_width, W = p->width, H = p->height;
    V2 tl = v2(fw, fw), tr = v2(W - fw, fw), br = v2(W - fw, H - fw), bl = v2(fw, H - fw);
    switch (wi) {
        case 0: *a = tl; *b = bl; break;
        case 1: *a = bl; *b = br; break;
        case 2: *a = tr; *b = br; break;
        default: *a = tl; *b = tr; break;
    }
}
static int circle_wall(const SapioParams *p, int wi, V2 c, double r, V2 *n_out, V2 *closest_out, double *dist_out) {
    V2 sa, sb;
    wall_segment(p, wi, &sa, &sb);
    V2 sd = vsub(sb, sa);
    double t = vdot(sd, vsub(c, sa)) / vdot(sd, sd);
    t = t < 0 ? 0 : (t > 1 ? 1 : t);
    V2 closest = vadd(sa, vmul(sd, t));
    double mind = r + p->frame_width;
    V2 delta = vsub(closest, c);
    double d2 = vdot(delta, delta);
    if (!(d2 < mind * mind)) return 0;
    double d = sqrt(d2);
    /* coincident: segment->tn = perp(normalize(b - a)) */
    V2 n;
    if (d != 0) n = vmul(delta, 1.0 / d);
    else { double L = sqrt(vdot(sd, sd)); n = vperp(vmul(sd, 1.0 / L)); }
    *n_out = n; *closest_out = closest; *dist_out = d;
    return 1;
}

static inline double k_scalar(const Body *a, const Body *b, V2 r1, V2 r2, V2 n) {
    double c1 = vcross(r1, n), c2 = vcross(r2, n);
    return (a->m_inv + a->i_inv * c1 * c1) + (b->m_inv + b->i_inv * c2 * c2);
}
static inline V2 rel_vel(const Body *a, const Body *b, V2 r1, V2 r2) {
    V2 v1 = vadd(a->v, vmul(vperp(r1), a->w));
    V2 v2_ = vadd(b->v, vmul(vperp(r2), b->w));
    return vsub(v2_, v1);
}
static inline void apply_impulses(Body *a, Body *b, V2 r1, V2 r2, V2 j) {
    a->v = vsub(a->v, vmul(j, a->m_inv)); a->w -= a->i_inv * vcross(r1, j);
    b->v = vadd(b->v, vmul(j, b->m_inv)); b->w += b->i_inv * vcross(r2, j);
}
static inline void apply_bias_impulses(Body *a, Body *b, V2 r1, V2 r2, V2 j) {
    a->vb = vsub(a->vb, vmul(j, a->m_inv)); a->wb -= a->i_inv * vcross(r1, j);
    b->vb = vadd(b->vb, vmul(j, b->m_inv)); b->wb += b->i_inv * vcross(r2, j);
}

/* cpArbiterPreStep (SURVEY A-4) */
static void contact_prestep(Contact *c, const Body *B, double dt, double slop, double bias_coef) {
    const Body *a = &B[c->a], *b = &B[c->b];
    c->nMass = 1.0 / k_scalar(a, b, c->r1, c->r2, c->n);
    c->tMass = 1.0 / k_scalar(a, b, c->r1, c->r2, vperp(c->n));
    double dist = vdot(vadd(vsub(c->r2, c->r1), vsub(b->p, a->p)), c->n);
    c->bias = -bias_coef * fmin(0.0, dist + slop) / dt;
    c->jBias = 0.0;
    c->bounce = vdot(rel_vel(a, b, c->r1, c->r2), c->n) * c->e;
}
/* cpArbiterApplyCachedImpulse */
static void contact_warm(Contact *c, Body *B, double dt_coef) {
    V2 j = vrotate(c->n, v2(c->jn, c->jt));
    apply_impulses(&B[c->a], &B[c->b], c->r1, c->r2, vmul(j, dt_coef));
}
/* cpArbiterApplyImpulse (SURVEY A-5) */
static void contact_apply(Contact *c, Body *B) {
    Body *a = &B[c->a], *b = &B[c->b];
    V2 n = c->n, r1 = c->r1, r2 = c->r2;
    V2 vb1 = vadd(a->vb, vmul(vperp(r1), a->wb));
    V2 vb2 = vadd(b->vb, vmul(vperp(r2), b->wb));
    V2 vr = rel_vel(a, b, r1, r2);
    double vbn = vdot(vsub(vb2, vb1), n);
    double vrn = vdot(vr, n);
    double vrt = vdot(vr, vperp(n));
    double jbn = (c->bias - vbn) * c->nMass;
    double jbnOld = c->jBias;
    c->jBias = fmax(jbnOld + jbn, 0.0);
    double jn = -(c->bounce + vrn) * c->nMass;
    double jnOld = c->jn;
    c->jn = fmax(jnOld + jn, 0.0);
    double jtMax = c->u * c->jn;
    double jt = -vrt * c->tMass;
    double jtOld = c->jt;
    double s = jtOld + jt;
    c->jt = s < -jtMax ? -jtMax : (s > jtMax ? jtMax : s);
    apply_bias_impulses(a, b, r1, r2, vmul(n, c->jBias - jbnOld));
    apply_impulses(a, b, r1, r2, vrotate(n, v2(c->jn - jnOld, c->jt - jtOld)));
}

/* cpPinJoint.c preStep / applyCachedImpulse / applyImpulse with both anchors at the centres (SURVEY A-6) */
static void pin_prestep(Pin *j, const Body *B, double rest, double dt, double bias_coef) {
    const Body *a = &B[j->a], *b = &B[j->b];
    V2 delta = vsub(b->p, a->p);
    double dist = sqrt(vdot(delta, delta));
    j->n = dist != 0 ? vmul(delta, 1.0 / dist) : v2(0, 0);
    j->nMass = 1.0 / (a->m_inv + b->m_inv);
    j->bias = -bias_coef * (dist - rest) / dt;     /* maxBias = INFINITY */
}
static void pin_warm(Pin *j, Body *B, double dt_coef) {
    V2 imp = vmul(j->n, j->jn * dt_coef);
    B[j->a].v = vsub(B[j->a].v, vmul(imp, B[j->a].m_inv));
    B[j->b].v = vadd(B[j->b].v, vmul(imp, B[j->b].m_inv));
}
static void pin_apply(Pin *j, Body *B) {
    Body *a = &B[j->a], *b = &B[j->b];
    double vrn = vdot(vsub(b->v, a->v), j->n);
    double jn = (j->bias - vrn) * j->nMass;        /* maxForce = INFINITY: no clamp */
    j->jn += jn;
    V2 imp = vmul(j->n, jn);
    a->v = vsub(a->v, vmul(imp, a->m_inv));
    b->v = vadd(b->v, vmul(imp, b->m_inv));
}

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* body table layout: [0, NC) cell rows, [NC, NC+ND) dead slots, [NC+ND, NC+ND+4) walls, then NC sensor bodies */
int oracle_space_step(SapioWorld *w) {
    const SapioParams *p = &w->p;
    const int NO = p->n_org, NC = NO * MAXC, ND = p->n_dead;
    const int WALL0 = NC + ND, SENS0 = WALL0 + SAPIO_NWALL, NB = SENS0 + NC;
    const float dtf = p->dt_phys;
    const double dt = (double)dtf;
    const int32_t tick = (int32_t)w->g_tick[0];
    int rc = 0;

    Body *B = (Body *)calloc((size_t)NB, sizeof(Body));
    uint8_t *solid = (uint8_t *)calloc((size_t)NB, 1);      /* 1 = takes part in collisions */
    uint8_t *fresh = (uint8_t *)calloc((size_t)NB, 1);

    /* (1) cpBodyUpdatePosition, MIRROR32: p = fma(v + v_bias, dt, p) in fp32, exactly what the GPU does */
    for (int o = 0; o < NO; o++) {
        if (w->org_state[o] != 1) continue;
        int born_now = (w->org_birth_tick[o] == tick);
        for (int k = 0; k < w->org_ncells[o]; k++) {
            int row = o * MAXC + k;
            if (w->c_alive[row] != SAPIO_CELL_ALIVE) continue;
            for (int c = 0; c < 2; c++) {
                float s = w->c_vel[2 * row + c] + w->c_vbias[2 * row + c];
                w->c_pos[2 * row + c] = fmaf(s, dtf, w->c_pos[2 * row + c]);
                w->c_vbias[2 * row + c] = 0;
            }
            { float s = w->c_angvel[row] + w->c_wbias[row]; w->c_ang[row] = fmaf(s, dtf, w->c_ang[row]); w->c_wbias[row] = 0; }
            Body *b = &B[row];
            double m = w->c_mass[row], r = w->c_size[row] * 0.5;
            b->p = v2(w->c_pos[2 * row], w->c_pos[2 * row + 1]);
            b->v = v2(w->c_vel[2 * row], w->c_vel[2 * row + 1]);
            b->w = w->c_angvel[row];
            b->m_inv = 1.0 / m; b->i_inv = 1.0 / (0.5 * m * r * r);      /* moment_for_circle (physics.py:112) */
            b->r = r; b->e = w->c_elast[row]; b->u = w->c_fric[row]; b->dynamic = 1;
            solid[row] = 1; fresh[row] = (uint8_t)born_now;
            int t = w->c_type[row];
            if (t == SAPIO_T_EYE || t == SAPIO_T_OLFACTORY) {
                for (int c = 0; c < 2; c++) {
                    float s = w->c_svel[2 * row + c] + w->c_svbias[2 * row + c];
                    w->c_spos[2 * row + c] = fmaf(s, dtf, w->c_spos[2 * row + c]);
                    w->c_svbias[2 * row + c] = 0;
                }
                Body *sb = &B[SENS0 + row];
                sb->p = v2(w->c_spos[2 * row], w->c_spos[2 * row + 1]);
                sb->v = v2(w->c_svel[2 * row], w->c_svel[2 * row + 1]);
                sb->m_inv = 1.0; sb->i_inv = 0.0;   /* mass 1 (sprites.py:303); no torque ever reaches it */
                sb->dynamic = 1;
            }
        }
    }
    for (int d = 0; d < ND; d++) {
        if (!w->d_state[d]) continue;
        for (int c = 0; c < 2; c++) {
            float s = w->d_vel[2 * d + c] + w->d_vbias[2 * d + c];
            w->d_pos[2 * d + c] = fmaf(s, dtf, w->d_pos[2 * d + c]);
            w->d_vbias[2 * d + c] = 0;
        }
        { float s = w->d_angvel[d] + w->d_wbias[d]; w->d_ang[d] = fmaf(s, dtf, w->d_ang[d]); w->d_wbias[d] = 0; }
        Body *b = &B[NC + d];
        double r = w->d_size[d] * 0.5;
        b->p = v2(w->d_pos[2 * d], w->d_pos[2 * d + 1]);
        b->v = v2(w->d_vel[2 * d], w->d_vel[2 * d + 1]);
        b->w = w->d_angvel[d];
        b->m_inv = w->d_mass[d] == 0 ? INFINITY : 1.0 / (double)w->d_mass[d];   /* cpBodySetMass */
        b->i_inv = 1.0 / (0.5 * (double)w->d_mass0[d] * r * r);                  /* _set_mass leaves the moment (sprites.py:1511) */
        b->r = r; b->e = w->d_elast[d]; b->u = w->d_fric[d]; b->dynamic = 1;
        solid[NC + d] = 1; fresh[NC + d] = (w->d_state[d] == 2);
    }
    for (int wi = 0; wi < SAPIO_NWALL; wi++) {
        Body *b = &B[WALL0 + wi];
        b->e = p->wall_elast; b->u = p->wall_fric; b->r = p->frame_width;   /* static: m_inv = i_inv = 0, p = 0 */
    }

    /* (2) collision detection. Candidate generation: uniform grid (any exact generator is equivalent). */
    const double gc = p->grid_cell;
    const int gx = p->grid_nx, gy = p->grid_ny;
    int nsolid = 0;
    for (int i = 0; i < WALL0; i++) nsolid += solid[i];
    int *cell_of = (int *)malloc(sizeof(int) * (size_t)(WALL0 > 0 ? WALL0 : 1));
    int *cstart = (int *)calloc((size_t)gx * gy + 1, sizeof(int));
    int *sorted = (int *)malloc(sizeof(int) * (size_t)(nsolid > 0 ? nsolid : 1));
    for (int i = 0; i < WALL0; i++) {
        if (!solid[i]) continue;
        int cx = (int)floor(B[i].p.x / gc) + 1, cy = (int)floor(B[i].p.y / gc) + 1;
        cx = cx < 0 ? 0 : (cx >= gx ? gx - 1 : cx); cy = cy < 0 ? 0 : (cy >= gy ? gy - 1 : cy);
        cell_of[i] = cy * gx + cx;
        cstart[cell_of[i] + 1]++;
    }
    for (int c = 0; c < gx * gy; c++) cstart[c + 1] += cstart[c];
    {
        int *fill = (int *)malloc(sizeof(int) * (size_t)gx * gy);
        memcpy(fill, cstart, sizeof(int) * (size_t)gx * gy);
        for (int i = 0; i < WALL0; i++) if (solid[i]) sorted[fill[cell_of[i]]++] = i;
        free(fill);
    }

    int xcap = p->x_cap, xn = 0, overflow = 0;
    uint64_t *xkey = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(xcap + 1));
    for (int i = 0; i < WALL0; i++) {
        if (!solid[i]) continue;
        int ci = cell_of[i], cx = ci % gx, cy = ci / gx;
        int org_i = i < NC ? i / MAXC : -1;
        for (int yy = cy - 1; yy <= cy + 1; yy++) for (int xx = cx - 1; xx <= cx + 1; xx++) {
            if (xx < 0 || yy < 0 || xx >= gx || yy >= gy) continue;
            for (int s = cstart[yy * gx + xx]; s < cstart[yy * gx + xx + 1]; s++) {
                int j = sorted[s];
                if (j <= i) continue;
                if (org_i >= 0 && j < NC && j / MAXC == org_i) continue;       /* intra pairs: handled per organism */
                double d2;
                if (!circles_overlap(B[i].p.x, B[i].p.y, B[i].r, B[j].p.x, B[j].p.y, B[j].r, &d2)) continue;
                if (xn < xcap) xkey[xn++] = ((uint64_t)(uint32_t)i << 32) | (uint32_t)j; else overflow = 1;
            }
        }
        if (i >= NC) {                                                           /* dead cell against the 4 walls */
            for (int wi = 0; wi < SAPIO_NWALL; wi++) {
                V2 n, cl; double d;
                if (circle_wall(p, wi, B[i].p, B[i].r, &n, &cl, &d)) {
                    if (xn < xcap) xkey[xn++] = ((uint64_t)(uint32_t)i << 32) | (uint32_t)(WALL0 + wi); else overflow = 1;
                }
            }
        }
    }
    qsort(xkey, (size_t)xn, sizeof(uint64_t), cmp_u64);

    /* contact construction helper */
    const double slop = p->collision_slop;
    const double bias_coef = 1.0 - pow((double)p->collision_bias, dt);
    const double jbias_coef = 1.0 - pow((double)p->joint_error_bias, dt);
    const double dt_coef = w->g_stepped[0] ? 1.0 : 0.0;      /* prev_dt == 0 on the very first step */

    Contact *X = (Contact *)calloc((size_t)(xn > 0 ? xn : 1), sizeof(Contact));
    const uint64_t *pk = w->x_key; const int pn = w->g_x_n[0];
    for (int c = 0; c < xn; c++) {
        int a = (int)(xkey[c] >> 32), b = (int)(xkey[c] & 0xFFFFFFFFu);
        Contact *ct = &X[c];
        ct->a = a; ct->b = b;
        if (b >= WALL0) {
            V2 n, cl; double d;
            circle_wall(p, b - WALL0, B[a].p, B[a].r, &n, &cl, &d);
            ct->n = n; ct->r1 = vmul(n, B[a].r); ct->r2 = vsub(cl, vmul(n, p->frame_width));
        } else {
            V2 delta = vsub(B[b].p, B[a].p);
            double d = sqrt(vdot(delta, delta));
            V2 n = d != 0 ? vmul(delta, 1.0 / d) : v2(1, 0);
            ct->n = n; ct->r1 = vmul(n, B[a].r); ct->r2 = vmul(n, -B[b].r);
        }
        ct->e = B[a].e * B[b].e; ct->u = B[a].u * B[b].u;
        ct->jn = ct->jt = 0;
        if (!fresh[a] && !(b < WALL0 && fresh[b]) && pn > 0) {                 /* persistent arbiter: keep jnAcc/jtAcc */
            const uint64_t *hit = (const uint64_t *)bsearch(&xkey[c], pk, (size_t)pn, sizeof(uint64_t), cmp_u64);
            if (hit) { ct->jn = w->x_jn[hit - pk]; ct->jt = w->x_jt[hit - pk]; }
        }
    }

    /* per-organism intra contacts and pins */
    Contact **IC = (Contact **)calloc((size_t)NO, sizeof(Contact *));
    int *icn = (int *)calloc((size_t)NO, sizeof(int));
    uint16_t **iccode = (uint16_t **)calloc((size_t)NO, sizeof(uint16_t *));
    Pin **PJ = (Pin **)calloc((size_t)NO, sizeof(Pin *));
    int *pjn = (int *)calloc((size_t)NO, sizeof(int));
    int **pjslot = (int **)calloc((size_t)NO, sizeof(int *));   /* >=0: pair index into j_jn, <0: -(row)-1 sensor pin */
    long n_ic_total = 0;
    for (int o = 0; o < NO; o++) {
        if (w->org_state[o] != 1) continue;
        int nc = w->org_ncells[o];
        IC[o] = (Contact *)calloc(SAPIO_ICAP, sizeof(Contact));
        iccode[o] = (uint16_t *)calloc(SAPIO_ICAP, sizeof(uint16_t));
        int n = 0;
        const uint16_t *pcode = w->ic_code + (size_t)o * SAPIO_ICAP;
        int pcn = w->org_ic_n[o];      /* a newborn's list is empty (ow_materialize), so no freshness test is needed here */
        for (int i = 0; i < nc; i++) {
            int ri = o * MAXC + i;
            if (!solid[ri]) continue;
            for (int j = i + 1; j < nc + SAPIO_NWALL; j++) {
                Contact ct; memset(&ct, 0, sizeof(ct));
                int code;
                if (j < nc) {
                    int rj = o * MAXC + j;
                    if (!solid[rj]) continue;
                    double d2;
                    if (!circles_overlap(B[ri].p.x, B[ri].p.y, B[ri].r, B[rj].p.x, B[rj].p.y, B[rj].r, &d2)) continue;
                    double d = sqrt(d2);
                    V2 delta = vsub(B[rj].p, B[ri].p);
                    V2 nn = d != 0 ? vmul(delta, 1.0 / d) : v2(1, 0);
                    ct.a = ri; ct.b = rj; ct.n = nn; ct.r1 = vmul(nn, B[ri].r); ct.r2 = vmul(nn, -B[rj].r);
                    code = i * 128 + j;
                } else {
                    int wi = j - nc;
                    V2 nn, cl; double d;
                    if (!circle_wall(p, wi, B[ri].p, B[ri].r, &nn, &cl, &d)) continue;
                    ct.a = ri; ct.b = WALL0 + wi; ct.n = nn; ct.r1 = vmul(nn, B[ri].r); ct.r2 = vsub(cl, vmul(nn, p->frame_width));
                    code = i * 128 + MAXC + wi;
                }
                ct.e = B[ct.a].e * B[ct.b].e; ct.u = B[ct.a].u * B[ct.b].u;
                for (int q = 0; q < pcn; q++) if (pcode[q] == code) { ct.jn = w->ic_jn[(size_t)o * SAPIO_ICAP + q]; ct.jt = w->ic_jt[(size_t)o * SAPIO_ICAP + q]; break; }
                if (n < SAPIO_ICAP) { IC[o][n] = ct; iccode[o][n] = (uint16_t)code; n++; } else overflow = 1;
            }
        }
        icn[o] = n; n_ic_total += n;
        /* constraints: sensor pins in row order, then all pairs (i<j) of living cells */
        int maxp = nc + nc * (nc - 1) / 2;
        PJ[o] = (Pin *)calloc((size_t)(maxp > 0 ? maxp : 1), sizeof(Pin));
        pjslot[o] = (int *)calloc((size_t)(maxp > 0 ? maxp : 1), sizeof(int));
        int m = 0;
        for (int k = 0; k < nc; k++) {
            int row = o * MAXC + k, t = w->c_type[row];
            if (!solid[row] || !(t == SAPIO_T_EYE || t == SAPIO_T_OLFACTORY)) continue;
            Pin *pj = &PJ[o][m]; pjslot[o][m] = -row - 1; m++;
            pj->a = SENS0 + row; pj->b = row;                                   /* PinJoint(newBody, newCell.body) (sprites.py:309,321) */
            pj->jn = w->c_spin_jn[row];
            pin_prestep(pj, B, 0.0, dt, jbias_coef);
        }
        for (int i = 0; i < nc; i++) {
            int ri = o * MAXC + i;
            if (!solid[ri]) continue;
            for (int j = i + 1; j < nc; j++) {
                int rj = o * MAXC + j;
                if (!solid[rj]) continue;
                Pin *pj = &PJ[o][m]; int q = sapio_pair_index(i, j); pjslot[o][m] = q; m++;
                pj->a = ri; pj->b = rj;
                pj->jn = w->j_jn[(size_t)o * SAPIO_NPAIR + q];
                /* rest length = centre distance when connectTo ran == |relative_pos_i - relative_pos_j| (sprites.py:1127-1133,1721) */
                double lx = (double)w->c_rel[2 * rj] - (double)w->c_rel[2 * ri], ly = (double)w->c_rel[2 * rj + 1] - (double)w->c_rel[2 * ri + 1];
                pin_prestep(pj, B, sqrt(lx * lx + ly * ly), dt, jbias_coef);
            }
        }
        pjn[o] = m;
    }

    /* (4) prestep arbiters (velocities are still pre-warm-start) */
    for (int c = 0; c < xn; c++) contact_prestep(&X[c], B, dt, slop, bias_coef);
    for (int o = 0; o < NO; o++) for (int c = 0; c < icn[o]; c++) contact_prestep(&IC[o][c], B, dt, slop, bias_coef);
    /* (5) cpBodyUpdateVelocity: gravity 0, damping 1, no forces => identity. */
    /* (6) warm start: arbiters, then constraints */
    for (int c = 0; c < xn; c++) contact_warm(&X[c], B, dt_coef);
    for (int o = 0; o < NO; o++) for (int c = 0; c < icn[o]; c++) contact_warm(&IC[o][c], B, dt_coef);
    for (int o = 0; o < NO; o++) for (int c = 0; c < pjn[o]; c++) pin_warm(&PJ[o][c], B, dt_coef);
    /* (7) iterations */
    for (int it = 0; it < p->iterations; it++) {
        for (int c = 0; c < xn; c++) contact_apply(&X[c], B);
        for (int o = 0; o < NO; o++) for (int c = 0; c < icn[o]; c++) contact_apply(&IC[o][c], B);
        for (int o = 0; o < NO; o++) for (int c = 0; c < pjn[o]; c++) pin_apply(&PJ[o][c], B);
    }

    /* write back */
    for (int o = 0; o < NO; o++) {
        if (w->org_state[o] != 1) continue;
        for (int k = 0; k < w->org_ncells[o]; k++) {
            int row = o * MAXC + k;
            if (!solid[row]) continue;
            const Body *b = &B[row];
            w->c_vel[2 * row] = (float)b->v.x; w->c_vel[2 * row + 1] = (float)b->v.y; w->c_angvel[row] = (float)b->w;
            w->c_vbias[2 * row] = (float)b->vb.x; w->c_vbias[2 * row + 1] = (float)b->vb.y; w->c_wbias[row] = (float)b->wb;
            int t = w->c_type[row];
            if (t == SAPIO_T_EYE || t == SAPIO_T_OLFACTORY) {
                const Body *sb = &B[SENS0 + row];
                w->c_svel[2 * row] = (float)sb->v.x; w->c_svel[2 * row + 1] = (float)sb->v.y;
                w->c_svbias[2 * row] = (float)sb->vb.x; w->c_svbias[2 * row + 1] = (float)sb->vb.y;
            }
        }
        for (int c = 0; c < icn[o]; c++) {
            w->ic_code[(size_t)o * SAPIO_ICAP + c] = iccode[o][c];
            w->ic_jn[(size_t)o * SAPIO_ICAP + c] = (float)IC[o][c].jn;
            w->ic_jt[(size_t)o * SAPIO_ICAP + c] = (float)IC[o][c].jt;
        }
        w->org_ic_n[o] = icn[o];
        for (int c = 0; c < pjn[o]; c++) {
            if (pjslot[o][c] < 0) w->c_spin_jn[-pjslot[o][c] - 1] = (float)PJ[o][c].jn;
            else w->j_jn[(size_t)o * SAPIO_NPAIR + pjslot[o][c]] = (float)PJ[o][c].jn;
        }
    }
    for (int d = 0; d < ND; d++) {
        if (!w->d_state[d]) continue;
        const Body *b = &B[NC + d];
        w->d_vel[2 * d] = (float)b->v.x; w->d_vel[2 * d + 1] = (float)b->v.y; w->d_angvel[d] = (float)b->w;
        w->d_vbias[2 * d] = (float)b->vb.x; w->d_vbias[2 * d + 1] = (float)b->vb.y; w->d_wbias[d] = (float)b->wb;
        w->d_state[d] = 1;
    }
    for (int c = 0; c < xn; c++) { w->x_key[c] = xkey[c]; w->x_jn[c] = (float)X[c].jn; w->x_jt[c] = (float)X[c].jt; }
    w->g_x_n[0] = xn;
    {   /* adjacency by body: partners ascending (smaller ids, then larger ids, walls last); walls own no list */
        int *cnt = (int *)calloc((size_t)WALL0 + 1, sizeof(int));
        for (int c = 0; c < xn; c++) { int a = (int)(xkey[c] >> 32), b = (int)(xkey[c] & 0xFFFFFFFFu); cnt[a]++; if (b < WALL0) cnt[b]++; }
        int acc = 0;
        for (int b = 0; b <= WALL0; b++) { w->xa_off[b] = acc; acc += (b < WALL0) ? cnt[b] : 0; }
        memset(cnt, 0, sizeof(int) * ((size_t)WALL0 + 1));
        for (int c = 0; c < xn; c++) { int a = (int)(xkey[c] >> 32), b = (int)(xkey[c] & 0xFFFFFFFFu); if (b < WALL0) w->xa_other[w->xa_off[b] + cnt[b]++] = (uint32_t)a; }
        for (int c = 0; c < xn; c++) { int a = (int)(xkey[c] >> 32), b = (int)(xkey[c] & 0xFFFFFFFFu); w->xa_other[w->xa_off[a] + cnt[a]++] = (uint32_t)b; }
        free(cnt);
    }
    w->g_stepped[0] = 1;
    w->g_stats[SAPIO_STAT_XCONTACTS] = xn;
    w->g_stats[SAPIO_STAT_ICONTACTS] = n_ic_total;
    if (overflow) { w->g_stats[SAPIO_STAT_OVERFLOW]++; rc = SAPIO_E_OVERFLOW; }

    for (int o = 0; o < NO; o++) { free(IC[o]); free(iccode[o]); free(PJ[o]); free(pjslot[o]); }
    free(IC); free(icn); free(iccode); free(PJ); free(pjn); free(pjslot);
    free(X); free(xkey); free(cell_of); free(cstart); free(sorted); free(B); free(solid); free(fresh);
    return rc;
}

/* Sprite.updateSprite (physics.py:250-275, SURVEY E-8): every solid sprite of env.sprites loses friction % of v and omega
 * per second of wall time; sensor bodies are not sprites and are never damped. */
int oracle_friction(SapioWorld *w) {
    const SapioParams *p = &w->p;
    double uDiff = p->dt_wall, f = p->friction_pct / 100.0;
    for (int o = 0; o < p->n_org; o++) {
        if (w->org_state[o] != 1) continue;
        for (int k = 0; k < w->org_ncells[o]; k++) {
            int row = o * MAXC + k;
            if (w->c_alive[row] != SAPIO_CELL_ALIVE) continue;
            for (int c = 0; c < 2; c++) { double v = w->c_vel[2 * row + c]; w->c_vel[2 * row + c] = (float)(v - f * v * uDiff); }
            double a = w->c_angvel[row]; w->c_angvel[row] = (float)(a - f * a * uDiff);
        }
    }
    for (int d = 0; d < p->n_dead; d++) {
        if (!w->d_state[d]) continue;
        for (int c = 0; c < 2; c++) { double v = w->d_vel[2 * d + c]; w->d_vel[2 * d + c] = (float)(v - f * v * uDiff); }
        double a = w->d_angvel[d]; w->d_angvel[d] = (float)(a - f * a * uDiff);
    }
    return 0;
}
