// orgsolve.cuh -- the per-organism part of Space.step: intra-organism contacts, sensor pins and the all-pairs pin joints of
// one organism, 10 Gauss-Seidel sweeps entirely in shared memory.
//
// Shape of the work (what bounds this kernel): per sweep an organism of n living cells has n(n-1)/2 pin joints that must be
// applied in lexicographic order. Pairs on one anti-diagonal t = i + j share no body and all their predecessors lie on
// earlier diagonals, so the order is kept exactly by walking 2n-3 wavefronts of at most n/2 pairs: a dependent chain of
// ~11 * (2n + contact levels) short steps per organism and tick. The kernel is therefore bound by the latency of that chain
// and by how many organisms are in flight per SM, not by memory:
//   * G = 4/8/16/32 lanes per organism (32/G organisms per warp, organisms of one warp have near-equal n)
//   * living cells are compacted (rank space): no liveness tests in the sweeps
//   * pair constants live in wavefront-major order: a wavefront reads one contiguous run of float4
//   * contacts are sorted by dependency level once per tick; a level is one step
//   * eight size classes (8..64 rows) so that the shared-memory footprint, which sets the occupancy, follows the size
#pragma once
#include "common.cuh"
#include "solver_api.h"
#ifndef JOINT_PREFETCH
#define JOINT_PREFETCH 1      /* pair record of the next wavefront fetched a step ahead */
#endif

// GC ("global constants"): the read-only part of a pin joint (nx, ny, bias) lives in a per-warp scratch in global memory (L2
// resident: it is rewritten for every organism the warp takes) and is fetched three wavefronts ahead; shared memory keeps only
// the accumulated impulse. 4 instead of 16 bytes per joint on chip: twice the organisms in flight for the large classes.
template <int CM, bool GC = false>
struct alignas(16) OrgS {
    static constexpr int NP = CM * (CM - 1) / 2;
    static constexpr int IC = 2 * CM < 16 ? 16 : 2 * CM;          // intra contacts kept on chip (global cap SAPIO_ICAP = 192)
    float4 pair[GC ? 1 : NP];        // wavefront-major: nx, ny, bias, jnAcc of a pin joint
    float jn[GC ? NP : 1];           // GC: jnAcc alone
    float4 vm[CM + 1];               // vx, vy, w, 1/m                       (index = rank among the living cells; [CM]: the walls' static body)
    float4 bb[CM + 1];               // bias velocity x, y, r/I, r
    float4 spin[CM];                 // sensor pin: nx, ny, bias, jnAcc
    // intra contacts: detected in canonical order (sorted code), then laid out in EXECUTION order (by dependency level, canonical
    // within a level) so that a level step reads three consecutive runs of float4
    float4 ce0[IC];                  // nx, ny, bias, bounce          (canonical order between detection and prestep_contacts)
    float4 ce1[IC];                  // 1 / k_n, 1 / k_t, mu, bits: canonical index | rank a << 8 | rank b << 16 (255: wall)
    float4 cacc[IC];                 // jnAcc, jtAcc, jBias, -
    float2 sv[CM];                   // sensor body velocity
    float2 p[CM], rel[CM];
    uint16_t ccode[IC];              // canonical key: row_i * 128 + row_j (walls: 64 + wall)
    uint8_t ca[IC], cbk[IC];         // ranks of the two cells (255: wall), canonical order
    uint8_t corder[IC], cpos[IC], clev[IC], sstart[IC + 2], lcnt[IC + 2];      // execution position -> canonical index and back; level steps
    uint8_t row_of[CM], last[CM], issens[CM];
    int n, nic, nsteps;
};

__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ int tri_f(int m) { return ((m + 1) >> 1) * ((m >> 1) + 1); }       // sum_{u=1..m} ceil(u/2)
// first slot of wavefront t (pairs i + j == t) for n cells
__device__ __forceinline__ int wf_off(int t, int n) { return t <= n ? tri_f(t - 1) : n * (n - 1) / 2 - tri_f(2 * n - 2 - t); }

// the cross contacts of the step as seen from a body (its adjacency), for the warm start of coupled organisms
struct CrossView { const int *off_all; const int *adj_c; const uint64_t *key; const float2 *n; const float *jn, *jt; int nx; };

__device__ __forceinline__ CrossView cross_view(const Workspace &W) {
    return CrossView{W.off_all, W.adj_c, W.nx_key, W.nx_n, W.nx_jn, W.nx_jt, W.flags[FL_NX]};
}

template <int CM, int G, bool GC = false>
struct OrgCtx {
    typedef OrgS<CM, GC> S;
    static constexpr int KC = (S::IC + G - 1) / G;                  // contacts per lane when all are touched
    static constexpr uint32_t GM = G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1u);
    S &s; const SapioWorld &w; const StepK &K;
    int o, nc, n, lane, sl, gshift;
    uint64_t sens;
    float4 *cst;                                                    // GC: this warp's constants scratch
    __device__ OrgCtx(S &s_, const SapioWorld &w_, const StepK &K_, int o_, float4 *cst_ = nullptr) : s(s_), w(w_), K(K_), o(o_), cst(cst_) {
        lane = lane_id(); sl = lane % G; gshift = lane - sl;
        nc = o >= 0 ? min(w.org_ncells[o], MAXC) : 0; n = 0; sens = 0;
    }
    __device__ uint32_t gballot(bool pred) const { return (__ballot_sync(FULL, pred) >> gshift) & GM; }
    __device__ static int wmax(int v) { for (int s_ = 16; s_; s_ >>= 1) v = max(v, __shfl_xor_sync(FULL, v, s_)); return v; }

    __device__ void load() {
        // rows in chunks of G: the living cells are compacted (rank = number of living cells in front), an organism of a class holds at
        // most CM of them whatever the rows it occupies (step.cuh: k_classify_count)
        int cnt = 0;
        const int ncw = wmax(nc);
        for (int h0 = 0; h0 < ncw; h0 += G) {
            const int k = h0 + sl, row = o * MAXC + k;
            const bool a = k < nc && w.c_alive[row] == SAPIO_CELL_ALIVE;
            const uint32_t am = gballot(a);
            const int r = cnt + __popc(am & ((1u << sl) - 1u));
            cnt += __popc(am);
            if (!a || r >= CM) continue;
            const bool sn = is_sensor_type(w.c_type[row]);
            const float m = w.c_mass[row], rad = w.c_size[row] * 0.5f;
            const float2 pc = reinterpret_cast<const float2 *>(w.c_pos)[row], v = reinterpret_cast<const float2 *>(w.c_vel)[row];
            const float2 vb = reinterpret_cast<const float2 *>(w.c_vbias)[row];
            const float minv = 1.0f / m, iinv = 1.0f / (0.5f * m * rad * rad);
            s.p[r] = pc; s.rel[r] = reinterpret_cast<const float2 *>(w.c_rel)[row];
            s.vm[r] = make_float4(v.x, v.y, w.c_angvel[row], minv);
            s.bb[r] = make_float4(vb.x, vb.y, iinv * rad, rad);
            s.row_of[r] = (uint8_t)k; s.issens[r] = sn;
            if (sn) {   // cpPinJoint preStep of the sensor pin (sensor body -> cell, anchors at the centres, rest length 0)
                const float2 ps = reinterpret_cast<const float2 *>(w.c_spos)[row];
                float dx = pc.x - ps.x, dy = pc.y - ps.y;
                float d = sqrtf(dx * dx + dy * dy);
                float inv = d != 0.f ? 1.0f / d : 0.f;
                s.spin[r] = make_float4(dx * inv, dy * inv, (float)(-K.bc_joint * (double)d / K.dt), w.c_spin_jn[row]);
                s.sv[r] = reinterpret_cast<const float2 *>(w.c_svel)[row];
            }
        }
        n = min(cnt, CM);
        if (cnt > CM && sl == 0) w.g_stats[SAPIO_STAT_OVERFLOW] = 1;                 // the classifier sized this organism wrongly: never silent
        if (sl == 0) { s.n = n; s.vm[CM] = make_float4(0.f, 0.f, 0.f, 0.f); s.bb[CM] = make_float4(0.f, 0.f, 0.f, w.p.frame_width); }   // a wall: infinite mass, radius = frame
        __syncwarp();
        uint32_t s0 = gballot(sl < n && s.issens[sl]), s1 = gballot(sl + G < n && s.issens[min(sl + G, CM - 1)]);
        sens = ((uint64_t)s1 << G) | s0;
    }

    // the context of an organism whose slab is still resident (components.cuh): what load() left in registers
    __device__ void resume() {
        n = s.n;
        uint32_t s0 = gballot(sl < n && s.issens[sl]), s1 = gballot(sl + G < n && s.issens[min(sl + G, CM - 1)]);
        sens = ((uint64_t)s1 << G) | s0;
    }

    // One pass over all pairs (i<j) of living cells in lexicographic order, a full group of lanes at a time: pin-joint
    // prestep (cpPinJoint.c preStep: n, bias; impulse of the previous step) and intra-organism collision detection
    // (cpCollision.c CircleToCircle) share the same separation, computed once in double from the fp32 positions.
    // Cell/wall contacts follow; the list is brought into canonical order (sorted code) when there are any.
    __device__ static void pair_of(int pidx, int n, int &i, int &j) {          // lexicographic pair index -> (i, j)
        const int tn = 2 * n - 1;
        i = (int)(((float)tn - sqrtf((float)(tn * tn - 8 * pidx))) * 0.5f);
        while (i > 0 && i * (tn - i) / 2 > pidx) i--;
        while ((i + 1) * (tn - i - 1) / 2 <= pidx) i++;
        j = pidx - i * (tn - i) / 2 + i + 1;
    }
    __device__ int slot_of(int i, int j) const { const int t = i + j; return wf_off(t, n) + i - max(0, t - (n - 1)); }
    // the lexicographic pairs a lane visits (sl, sl + G, sl + 2G, ...) as (i, j) kept incrementally: row i holds j = i + 1 .. n - 1
    __device__ static void pair_norm(int n, int &i, int &j) { while (j >= n && i < n - 1) { j += i + 2 - n; i++; } }
    __device__ void pair_first(int &i, int &j) const { i = 0; j = 1 + sl; pair_norm(n, i, j); }
    __device__ void pair_next(int &i, int &j) const { j += G; pair_norm(n, i, j); }       // valid while i < n - 1

    __device__ void prestep_pairs() {
        const int NPn = n * (n - 1) / 2, NPmax = wmax(NPn);
        int nic = 0; bool over = false;
        const float *gj = w.j_jn + (size_t)max(o, 0) * SAPIO_NPAIR;
        // one pass: the cached impulse of a pair is requested first (a gather from global memory) and arrives under the double-precision
        // geometry of the same pair; constants and impulse go to the pair's wavefront slot as one float4
        int i, j; pair_first(i, j);
        for (int base = 0; base < NPmax; base += G, pair_next(i, j)) {
            bool hit = false; float nxf = 0.f, nyf = 0.f, bias_c = 0.f;
            if (i < n - 1) {
                const float jn_cached = gj[pair_index(s.row_of[i], s.row_of[j])];
                const float2 pi = s.p[i], pj = s.p[j], ri = s.rel[i], rj = s.rel[j];
                const float radi = s.bb[i].w, radj = s.bb[j].w;
                // Squared lengths exactly in double (the contact predicate below is bit-exact against the oracle). The two lengths and the
                // reciprocal are fp32 hardware approximations refined by ONE Newton step in double (error 1e-7 squared), so unit vector and
                // bias come out as the correctly rounded fp32 values the double-precision sqrt / divide gave -- for 11 double-pipe
                // instructions per pin instead of ~100 (four library routines that used to be 15 % of the kernel).
                double dx = __dsub_rn((double)pj.x, (double)pi.x), dy = __dsub_rn((double)pj.y, (double)pi.y);
                double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                double lx = (double)rj.x - (double)ri.x, ly = (double)rj.y - (double)ri.y;
                double r2 = lx * lx + ly * ly;                       // |relative_pos_i - relative_pos_j|^2 (connectTo at spawn)
                const float d2f = (float)d2, r2f = (float)r2;
                double d = 0.0, rest = 0.0;
#ifdef SAPIO_DEBUG_FLAGS                               /* A/B switch 2: the library sqrt / divide of round 1 */
                if (K.dbg & 2) {
                    d = sqrt(d2); rest = sqrt(r2);
                    if (d != 0.0) { const double inv = 1.0 / d; nxf = (float)(dx * inv); nyf = (float)(dy * inv); }
                } else
#endif
                if (d2f > 0.f) {
                    const float iv = rsqrtf(d2f);
                    d = (double)(d2f * iv); d = fma(fma(-d, d, d2), (double)(0.5f * iv), d);                 // sqrt(d2)
                    double y = (double)iv; y = fma(y, fma(-d, y, 1.0), y);                                   // 1 / d
                    nxf = (float)(dx * y); nyf = (float)(dy * y);
                }
                if (r2f > 0.f && !(K.dbg & 2)) { const float iv = rsqrtf(r2f); rest = (double)(r2f * iv); rest = fma(fma(-rest, rest, r2), (double)(0.5f * iv), rest); }
                float bias = (float)(-(K.bc_joint / K.dt) * (d - rest));
                const int q = slot_of(i, j);
                if (GC) { cst[q] = make_float4(nxf, nyf, bias, 0.f); s.jn[q] = jn_cached; }
                else s.pair[q] = make_float4(nxf, nyf, bias, jn_cached);
                double md = __dadd_rn((double)radi, (double)radj);
                hit = d2 < __dmul_rn(md, md);
                if (hit) { if (d2 == 0.0) { nxf = 1.f; nyf = 0.f; } bias_c = contact_bias(d, radi, radj, w.p, K); }   // coincident centres: cpv(1, 0)
            }
            const uint32_t m = gballot(hit);
            if (hit) {
                int slot = nic + __popc(m & ((1u << sl) - 1u));
                if (slot < S::IC) {
                    s.ccode[slot] = (uint16_t)(s.row_of[i] * 128 + s.row_of[j]); s.ca[slot] = (uint8_t)i; s.cbk[slot] = (uint8_t)j;
                    s.ce0[slot] = make_float4(nxf, nyf, bias_c, 0.f);
                } else over = true;
            }
            nic += __popc(m);
        }
        // walls
        bool anywall = false;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int r = sl + G * h;
            bool near = false; float2 pi = make_float2(0.f, 0.f); float radi = 0.f;
            if (r < n) {
                pi = s.p[r]; radi = s.bb[r].w;
                float fw2 = 2.0f * w.p.frame_width + radi + 1.0f;                    // cheap reject: not near any border
                near = pi.x < fw2 || pi.y < fw2 || pi.x > w.p.width - fw2 || pi.y > w.p.height - fw2;
            }
            if (!__any_sync(FULL, near)) continue;
            for (int wl = 0; wl < SAPIO_NWALL; wl++) {
                WallHit hw; bool hit = near && circle_wall(w.p, wl, pi.x, pi.y, radi, &hw);
                const uint32_t m = gballot(hit);
                if (hit) {
                    int slot = nic + __popc(m & ((1u << sl) - 1u));
                    if (slot < S::IC) {
                        s.ccode[slot] = (uint16_t)(s.row_of[r] * 128 + MAXC + wl); s.ca[slot] = (uint8_t)r; s.cbk[slot] = 255;
                        s.ce0[slot] = make_float4(hw.nx, hw.ny, contact_bias(hw.d, radi, w.p.frame_width, w.p, K), 0.f);
                    } else over = true;
                    anywall = true;
                }
                nic += __popc(m);
            }
        }
        if (gballot(over) || nic > S::IC) { nic = S::IC; if (sl == 0) w.g_stats[SAPIO_STAT_OVERFLOW] = 1; }
        if (sl == 0) s.nic = nic;
        __syncwarp();
        if (__any_sync(FULL, anywall)) {       // merge the wall contacts into canonical order: rank = number of smaller codes
            uint16_t code[KC]; uint8_t a_[KC], b_[KC]; float4 g_[KC]; int rk[KC];
#pragma unroll
            for (int q = 0; q < KC; q++) {
                const int c = sl + G * q;
                rk[q] = -1;
                if (c < nic) {
                    code[q] = s.ccode[c]; a_[q] = s.ca[c]; b_[q] = s.cbk[c]; g_[q] = s.ce0[c];
                    int r = 0;
                    for (int e = 0; e < nic; e++) r += s.ccode[e] < code[q];
                    rk[q] = r;
                }
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < KC; q++)
                if (rk[q] >= 0) { const int r = rk[q]; s.ccode[r] = code[q]; s.ca[r] = a_[q]; s.cbk[r] = b_[q]; s.ce0[r] = g_[q]; }
            __syncwarp();
        }
    }

    // dependency levels: contact c runs after every earlier contact (list order) that shares a cell with it. The contacts are
    // then ordered by level (stable) and cut into steps of at most G contacts of one level.
    __device__ void schedule() {
        if (sl == 0) {
            const int nic = s.nic;
            for (int k = 0; k < n; k++) s.last[k] = 0;
            for (int L = 0; L <= nic + 1 && L < S::IC + 2; L++) s.lcnt[L] = 0;
            int mx = 0;
            for (int c = 0; c < nic; c++) {
                const int a = s.ca[c], b = s.cbk[c];
                int L = s.last[a];
                if (b != 255 && s.last[b] > L) L = s.last[b];
                L++;
                s.clev[c] = (uint8_t)L; s.last[a] = (uint8_t)L;
                if (b != 255) s.last[b] = (uint8_t)L;
                s.lcnt[L]++;
                mx = max(mx, L);
            }
            int ns = 0, pos = 0;
            for (int L = 1; L <= mx; L++) {                     // lcnt[L] becomes the write cursor of level L
                const int cnt = s.lcnt[L];
                for (int k = 0; k < cnt; k += G) s.sstart[ns++] = (uint8_t)(pos + k);
                s.lcnt[L] = (uint8_t)pos; pos += cnt;
            }
            s.sstart[ns] = (uint8_t)nic;
            for (int c = 0; c < nic; c++) { const int ps = s.lcnt[s.clev[c]]++; s.corder[ps] = (uint8_t)c; s.cpos[c] = (uint8_t)ps; }
            s.nsteps = ns;
        }
        __syncwarp();
    }

    // cpArbiterPreStep, and the move from canonical to execution order: effective masses, friction, and (FRESH) the persistent
    // arbiter's jnAcc / jtAcc of the previous step's intra list (sorted by code: binary search) and the bounce from the velocities
    // before any cached impulse. !FRESH (a later pass of the cooperative kernel): impulses and bounce come from load_contact_state.
    template <bool FRESH>
    __device__ void prestep_contacts() {
        const int nic = s.nic;
        const int pn = (FRESH && o >= 0) ? w.org_ic_n[o] : 0;
        const uint16_t *pc = w.ic_code + (size_t)max(o, 0) * SAPIO_ICAP;
        float4 gq[KC];                                               // the canonical geometry leaves ce0 before the execution-ordered records arrive
#pragma unroll
        for (int q = 0; q < KC; q++) { const int c = sl + G * q; if (c < nic) gq[q] = s.ce0[c]; }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < KC; q++) {
            const int c = sl + G * q;
            if (c < nic) {
                const int a = s.ca[c], b = s.cbk[c];
                const float4 g = gq[q];
                const int rowa = o * MAXC + s.row_of[a];
                const float4 va = s.vm[a], ba = s.bb[a];
                float4 vb = make_float4(0.f, 0.f, 0.f, 0.f), bbv = make_float4(0.f, 0.f, 0.f, w.p.frame_width);
                float e = w.c_elast[rowa] * w.p.wall_elast, mu = w.c_fric[rowa] * w.p.wall_fric;
                if (b != 255) { const int rowb = o * MAXC + s.row_of[b]; vb = s.vm[b]; bbv = s.bb[b]; e = w.c_elast[rowa] * w.c_elast[rowb]; mu = w.c_fric[rowa] * w.c_fric[rowb]; }
                float jn = 0.f, jt = 0.f, bo = 0.f;
                if (FRESH) {
                    const uint16_t code = s.ccode[c];
                    int lo = 0, hi = pn;
                    while (lo < hi) { int mid = (lo + hi) >> 1; if (pc[mid] < code) lo = mid + 1; else hi = mid; }
                    if (lo < pn && pc[lo] == code) { jn = w.ic_jn[(size_t)o * SAPIO_ICAP + lo]; jt = w.ic_jt[(size_t)o * SAPIO_ICAP + lo]; }
                    bo = ((vb.x - va.x) * g.x + (vb.y - va.y) * g.y) * e;
                }
                const int ps = s.cpos[c];
                s.ce0[ps] = make_float4(g.x, g.y, g.z, bo);
                s.ce1[ps] = make_float4(1.0f / (va.w + vb.w), 1.0f / (va.w + vb.w + ba.z * ba.w + bbv.z * bbv.w), mu, __int_as_float(c | (a << 8) | ((b == 255 ? CM : b) << 16)));
                s.cacc[ps] = make_float4(jn, jt, 0.f, 0.f);
            }
        }
        __syncwarp();
    }

    // one sweep over the intra contacts in canonical order (level steps): cpArbiterApplyImpulse (SURVEY A-5), circles: r x n = 0.
    // The records of step st + 1 are fetched while step st computes (nobody else writes them); a wall is the static body in slot CM,
    // whose state the arithmetic leaves unchanged (1/m = r/I = 0), so the step has no branch on the kind of contact.
    template <bool WARM>
    __device__ void contacts_pass(float dt_coef, int nsteps_max) {
        static_assert(!WARM, "cached contact impulses are applied by warm_start");
        const int ns = s.nsteps;
        int s1 = s.sstart[1], s2 = s.sstart[2];
        int pos = s.sstart[0] + sl;
        bool act = 0 < ns && pos < s1;
        float4 e0 = make_float4(0.f, 0.f, 0.f, 0.f), e1 = e0, ac = e0;
        if (act) { e0 = s.ce0[pos]; e1 = s.ce1[pos]; ac = s.cacc[pos]; }
        for (int st = 0; st < nsteps_max; st++) {
            const int s3 = s.sstart[min(st + 3, S::IC + 1)];
            const int posn = s1 + sl;
            const bool actn = st + 1 < ns && posn < s2;
            float4 e0n = make_float4(0.f, 0.f, 0.f, 0.f), e1n = e0n, acn = e0n;
            if (actn) { e0n = s.ce0[posn]; e1n = s.ce1[posn]; acn = s.cacc[posn]; }
            if (act) {
                const int desc = __float_as_int(e1.w), a = (desc >> 8) & 255, b = desc >> 16;
                float4 va = s.vm[a], ba = s.bb[a], vb = s.vm[b], bbv = s.bb[b];
                const float nx = e0.x, ny = e0.y;
                const float nMass = e1.x, tMass = e1.y;
                const float dvx = vb.x - va.x, dvy = vb.y - va.y;
                const float vbn = (bbv.x - ba.x) * nx + (bbv.y - ba.y) * ny;
                const float vrn = dvx * nx + dvy * ny;
                const float vrt = (dvy * nx - dvx * ny) - bbv.w * vb.z - ba.w * va.z;
                const float jbnOld = ac.z, jBias = fmaxf(jbnOld + (e0.z - vbn) * nMass, 0.0f);
                float dn, dtg, jnS, jtS;
                contact_accumulate(ac.x, ac.y, -(e0.w + vrn) * nMass, -vrt * tMass, e1.z, jnS, jtS, dn, dtg);       // steptypes.cuh: double accuracy in fp32
                s.cacc[pos] = make_float4(jnS, jtS, jBias, 0.f);
                const float bj = jBias - jbnOld;
                ba.x -= nx * bj * va.w; ba.y -= ny * bj * va.w;
                bbv.x += nx * bj * vb.w; bbv.y += ny * bj * vb.w;
                const float jx = nx * dn - ny * dtg, jy = nx * dtg + ny * dn;
                va.x -= jx * va.w; va.y -= jy * va.w; va.z -= ba.z * dtg;
                vb.x += jx * vb.w; vb.y += jy * vb.w; vb.z -= bbv.z * dtg;
                reinterpret_cast<float2 *>(&s.bb[a])[0] = make_float2(ba.x, ba.y);
                s.vm[a] = va;
                reinterpret_cast<float2 *>(&s.bb[b])[0] = make_float2(bbv.x, bbv.y);
                s.vm[b] = vb;
            }
            pos = posn; act = actn; e0 = e0n; e1 = e1n; ac = acn;
            s1 = s2; s2 = s3;
            __syncwarp();
        }
        (void)dt_coef;
    }

    // Warm start (cpArbiterApplyCachedImpulse, cached pin impulses), exact. Cached impulses do not depend on the velocities, so
    // their order is immaterial -- and they must not be applied one by one in fp32: the n(n-1)/2 pins of an organism are redundant,
    // their cached impulses drift in the null space (1e4..1e5 after a few hundred ticks) and cancel to a few px/s; a dead cell
    // wedged between pinned cells carries contact impulses of 1e5. Every lane sums, in double, everything its cells receive
    // (pins, intra contacts, the sensor pin, and for a coupled organism its cross contacts) and applies the sum once.
    __device__ void warm_start(float dt_coef, const CrossView *X) {
        const double dt = (double)dt_coef;
        // pins of the lane's two cells in one loop: two independent chains of loads and double FMAs
        double pax[2] = {0.0, 0.0}, pay[2] = {0.0, 0.0};
        {
            const int r0 = sl, r1 = sl + G;
            const bool v1 = r1 < n;
            if (r0 < n) {
                int off0 = r0 >= 1 ? wf_off(r0, n) : 0, off1 = v1 ? wf_off(r1, n) : 0;     // first slot of wavefront t = r + j, kept incrementally
#pragma unroll 4
                for (int j = 0; j < n; j++) {
                    {
                        const int t = r0 + j, imin = max(0, t - (n - 1));
                        if (j != r0) {
                            const int q = off0 + min(r0, j) - imin;
                            float nx, ny, jn;
                            if (GC) { const float4 c4 = __ldcg(&cst[q]); nx = c4.x; ny = c4.y; jn = s.jn[q]; }
                            else { const float4 pr = s.pair[q]; nx = pr.x; ny = pr.y; jn = pr.w; }
                            const double imp = (r0 < j ? -dt : dt) * (double)jn;                 // body a of a pin gets -n j, body b +n j
                            pax[0] = fma((double)nx, imp, pax[0]); pay[0] = fma((double)ny, imp, pay[0]);
                        }
                        off0 += max(((t - 1) >> 1) - imin + 1, 0);
                    }
                    if (v1) {
                        const int t = r1 + j, imin = max(0, t - (n - 1));
                        if (j != r1) {
                            const int q = off1 + min(r1, j) - imin;
                            float nx, ny, jn;
                            if (GC) { const float4 c4 = __ldcg(&cst[q]); nx = c4.x; ny = c4.y; jn = s.jn[q]; }
                            else { const float4 pr = s.pair[q]; nx = pr.x; ny = pr.y; jn = pr.w; }
                            const double imp = (r1 < j ? -dt : dt) * (double)jn;
                            pax[1] = fma((double)nx, imp, pax[1]); pay[1] = fma((double)ny, imp, pay[1]);
                        }
                        off1 += max(((t - 1) >> 1) - imin + 1, 0);
                    }
                }
            }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int r = sl + G * h;
            if (r < n) {
                double ax = pax[h], ay = pay[h], at = 0.0;         // linear impulse received; tangential impulses (both bodies: w -= r/I * jt)
                const int nic = s.nic;
                for (int c = 0; c < nic; c++) {                    // execution order (any order does: the sums are exact to 1e-16)
                    const int desc = __float_as_int(s.ce1[c].w), a = (desc >> 8) & 255, b = desc >> 16;
                    if (a != r && b != r) continue;
                    const float4 e0 = s.ce0[c], ac = s.cacc[c];
                    const double jn = (double)ac.x * dt, jt = (double)ac.y * dt, nx = (double)e0.x, ny = (double)e0.y;
                    const double jx = nx * jn - ny * jt, jy = nx * jt + ny * jn;
                    if (a == r) { ax -= jx; ay -= jy; } else { ax += jx; ay += jy; }
                    at += jt;
                }
                if (X) {
                    const uint32_t row = (uint32_t)(o * MAXC + s.row_of[r]);
                    for (int q = X->off_all[row], e = X->off_all[row + 1]; q < e; q++) {
                        const int c = X->adj_c[q];
                        if (c < 0 || c >= X->nx) continue;
                        const float2 nn = X->n[c];
                        const double jn = (double)X->jn[c] * dt, jt = (double)X->jt[c] * dt, nx = (double)nn.x, ny = (double)nn.y;
                        const double jx = nx * jn - ny * jt, jy = nx * jt + ny * jn;
                        if ((uint32_t)(X->key[c] >> 32) == row) { ax -= jx; ay -= jy; } else { ax += jx; ay += jy; }
                        at += jt;
                    }
                }
                if ((sens >> r) & 1ull) {                            // sensor pin: the cell is its body b, the sensor body (mass 1) its body a
                    const float4 sp = s.spin[r];
                    const double jn = (double)sp.w * dt;
                    ax = fma((double)sp.x, jn, ax); ay = fma((double)sp.y, jn, ay);
                    const float2 vs = s.sv[r];
                    s.sv[r] = make_float2((float)((double)vs.x - (double)sp.x * jn), (float)((double)vs.y - (double)sp.y * jn));
                }
                const float4 v = s.vm[r];
                s.vm[r] = make_float4((float)fma(ax, (double)v.w, (double)v.x), (float)fma(ay, (double)v.w, (double)v.y),
                                      (float)((double)v.z - (double)s.bb[r].z * at), v.w);
            }
        }
        __syncwarp();
        zero_pin_acc();
    }
    // from here on the pin accumulators hold what this step adds (store_vel adds it to the cached impulse in global memory)
    __device__ void zero_pin_acc() {
        const int NPn = n * (n - 1) / 2;
        for (int q = sl; q < NPn; q += G) { if (GC) s.jn[q] = 0.f; else s.pair[q].w = 0.f; }
        __syncwarp();
    }

    // cpPinJoint.c with anchors at the centres: sensor pins (sensor body -> cell, rest 0), then all pairs (i<j) in
    // lexicographic order evaluated as anti-diagonals t = i + j
    template <bool WARM>
    __device__ void joints_pass(float dt_coef, int tmax) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int k = sl + G * h;
            if (k < n && ((sens >> k) & 1ull)) {
                const float4 sp = s.spin[k], vc = s.vm[k]; const float2 vs = s.sv[k];
                float jnv;
                if (WARM) jnv = sp.w * dt_coef;
                else {
                    jnv = fmaf(-(vc.x - vs.x), sp.x, fmaf(-(vc.y - vs.y), sp.y, sp.z)) * (1.0f / (1.0f + vc.w));      // (bias - v_rel.n) * k
                    s.spin[k].w = sp.w + jnv;
                }
                const float jx = sp.x * jnv, jy = sp.y * jnv;
                s.sv[k] = make_float2(vs.x - jx, vs.y - jy);                                   // sensor body: mass 1 (sprites.py:303)
                reinterpret_cast<float2 *>(&s.vm[k])[0] = make_float2(fmaf(jx, vc.w, vc.x), fmaf(jy, vc.w, vc.y));
            }
        }
        __syncwarp();
        // Per wavefront: barrier -> pair constants and two velocity loads -> (bias - v_rel.n) with the cancellation kept exact by
        // FMAs -> times the effective mass (hardware reciprocal: 1 ulp, multiplicative, so the error vanishes with the impulse) ->
        // two stores -> barrier.
        if (GC) { joints_wavefronts_gc<WARM>(dt_coef, tmax); return; }
        // Cursors instead of indices (everything in bytes, 16 per float4): in the first half of a sweep (t <= n - 1) a lane keeps its row
        // i = sl and walks j = t - sl, in the second half it keeps its column and walks the rows; imin = max(t - (n - 1), 0) covers both,
        // "this lane has a pair" is i < j, and the wavefront-major pair cursor advances by the width h + 1 - imin with h = (t - 1) / 2,
        // which is the same number for the odd and the even step of a double step. ~10 integer instructions per step, none of them on
        // the chain load -> impulse -> store -> barrier -> load.
        // The body is branch-free (a lane without a pair computes on garbage and stores nothing): a predicate-dependent branch
        // costs 12-25 cycles of a ~100-cycle step, and a straight-line body lets the next step's cursors be computed under the loads.
        char *const vmb = reinterpret_cast<char *>(s.vm), *const pqb = reinterpret_cast<char *>(s.pair);
        const int S16 = sl * 16, N16 = (n - 1) * 16;
        int T16 = 16, Q16 = S16, H16 = 16;                      // 16 t; 16 (off(t) + sl); 16 (h + 1)
        int M = max(T16 - N16, 0), I = S16 + M, J = T16 - I;
        bool act = I < J;
        float4 *pp = reinterpret_cast<float4 *>(pqb + Q16);
        float4 pr = *pp;                                          // constants and accumulator of the pair about to run: fetched a step ahead
        for (int t = 1; t <= tmax; t += 2, H16 += 16) {
#pragma unroll
            for (int u = 0; u < 2; u++) {
                // an idle lane reads whatever lies at its cursors (always inside the CTA's shared memory: |J| < 16 G < the size of the pair
                // table in front of vm, I and Q run at most 32 entries past their arrays) and stores nothing
                float2 *const pi = reinterpret_cast<float2 *>(vmb + I), *const pj = reinterpret_cast<float2 *>(vmb + J);
                const float4 vi = *reinterpret_cast<const float4 *>(pi), vj = *reinterpret_cast<const float4 *>(pj);
                const bool act_now = act;
                float4 *const pp_now = pp;
                Q16 += max(H16 - M, 0);                          // the next step's cursors, under the latency of the loads
                T16 += 16;
                M = max(T16 - N16, 0); I = S16 + M; J = T16 - I; act = I < J;
                pp = reinterpret_cast<float4 *>(pqb + Q16);
#if JOINT_PREFETCH
                const float4 prn = *pp;                          // nobody writes the next pair's record during this step
#endif
                float jnv;
                if (WARM) jnv = pr.w * dt_coef;
                else jnv = fmaf(-(vj.x - vi.x), pr.x, fmaf(-(vj.y - vi.y), pr.y, pr.z)) * rcp_approx(vi.w + vj.w);
                const float jx = pr.x * jnv, jy = pr.y * jnv;
                const float2 ni = make_float2(fmaf(-jx, vi.w, vi.x), fmaf(-jy, vi.w, vi.y)), nj = make_float2(fmaf(jx, vj.w, vj.x), fmaf(jy, vj.w, vj.y));
                if (act_now) {
                    if (!WARM) reinterpret_cast<float *>(pp_now)[3] = pr.w + jnv;
                    *pi = ni;
                    *pj = nj;
                }
                __syncwarp();
#if JOINT_PREFETCH
                pr = prn;
#else
                pr = *pp;
#endif
            }
        }
    }

    // GC variant of the wavefront loop: the constants of wavefront t + 3 are requested (L2) while wavefront t runs
    __device__ float4 fetch_cst(int &tp, int &offp) const {
        const int imin = max(0, tp - (n - 1)), imax = (tp - 1) >> 1;
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (imin + sl <= imax) c = __ldcg(&cst[offp + sl]);
        offp += max(imax - imin + 1, 0); tp++;
        return c;
    }
    template <bool WARM>
    __device__ void wavefront_gc(const float4 pr, int t, int &off, float dt_coef) {
        const int imin = max(0, t - (n - 1)), imax = (t - 1) >> 1, i = imin + sl;
        if (i <= imax) {
            const int j = t - i, q = off + sl;
            const float4 vi = s.vm[i], vj = s.vm[j];
            const float jold = s.jn[q];
            float jnv;
            if (WARM) jnv = jold * dt_coef;
            else {
                jnv = fmaf(-(vj.x - vi.x), pr.x, fmaf(-(vj.y - vi.y), pr.y, pr.z)) * rcp_approx(vi.w + vj.w);
                s.jn[q] = jold + jnv;
            }
            const float jx = pr.x * jnv, jy = pr.y * jnv;
            reinterpret_cast<float2 *>(&s.vm[i])[0] = make_float2(fmaf(-jx, vi.w, vi.x), fmaf(-jy, vi.w, vi.y));
            reinterpret_cast<float2 *>(&s.vm[j])[0] = make_float2(fmaf(jx, vj.w, vj.x), fmaf(jy, vj.w, vj.y));
        }
        off += max(imax - imin + 1, 0);
        __syncwarp();
    }
    template <bool WARM>
    __device__ void joints_wavefronts_gc(float dt_coef, int tmax) {
        constexpr int D = 8;                                   // prefetch distance in wavefronts (~100 cycles each) against L2 latency
        int tp = 1, offp = 0, off = 0;
        float4 c[D];
#pragma unroll
        for (int u = 0; u < D; u++) c[u] = fetch_cst(tp, offp);
        for (int t = 1; t <= tmax; t += D) {
#pragma unroll
            for (int u = 0; u < D; u++)
                if (t + u <= tmax) { wavefront_gc<WARM>(c[u], t + u, off, dt_coef); c[u] = fetch_cst(tp, offp); }
        }
    }

    __device__ void store_vel(bool pins = true) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int r = sl + G * h;
            if (r < n) {
                const int row = o * MAXC + s.row_of[r];
                const float4 v = s.vm[r], b = s.bb[r];
                reinterpret_cast<float2 *>(w.c_vel)[row] = make_float2(v.x, v.y); w.c_angvel[row] = v.z;
                reinterpret_cast<float2 *>(w.c_vbias)[row] = make_float2(b.x, b.y);
                if ((sens >> r) & 1ull) { reinterpret_cast<float2 *>(w.c_svel)[row] = s.sv[r]; w.c_spin_jn[row] = s.spin[r].w; }
            }
        }
        if (!pins) return;
        float *gj = w.j_jn + (size_t)max(o, 0) * SAPIO_NPAIR;
        const int NPn = n * (n - 1) / 2;
        int i, j; pair_first(i, j);
        for (int base = 0; base < NPn; base += 4 * G) {          // cached impulse + what this step added; four loads in flight per lane
            float *dst[4]; float old[4], add[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                dst[u] = nullptr; old[u] = 0.f; add[u] = 0.f;
                if (i < n - 1) {
                    dst[u] = &gj[pair_index(s.row_of[i], s.row_of[j])]; old[u] = *dst[u];
                    add[u] = GC ? s.jn[slot_of(i, j)] : s.pair[slot_of(i, j)].w;
                }
                pair_next(i, j);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) if (dst[u]) *dst[u] = (float)((double)old[u] + (double)add[u]);
        }
    }
    __device__ void store_contacts(float *ws_bounce, float *ws_jbias) {
        for (int c = sl; c < s.nic; c += G) {
            size_t g = (size_t)o * SAPIO_ICAP + c;
            const int pos = s.cpos[c];
            const float4 ac = s.cacc[pos];
            w.ic_code[g] = s.ccode[c]; w.ic_jn[g] = ac.x; w.ic_jt[g] = ac.y;
            if (ws_bounce) { ws_bounce[g] = s.ce0[pos].w; ws_jbias[g] = ac.z; }
        }
        if (sl == 0 && o >= 0) w.org_ic_n[o] = s.nic;
    }
    // coupled path: what prestep_pairs / schedule / prestep_contacts produced is parked in global memory (same struct layout) so
    // that the passes between the grid barriers do not detect and pre-step again
    __device__ void save_static(S *g) const {
        const int NPn = n * (n - 1) / 2, nic = s.nic;
        for (int q = sl; q < NPn; q += G) g->pair[q] = s.pair[q];
        for (int c = sl; c < nic; c += G) { g->ce0[c] = s.ce0[c]; g->ce1[c] = s.ce1[c]; g->ccode[c] = s.ccode[c]; g->cpos[c] = s.cpos[c]; }
        for (int q = sl; q <= s.nsteps; q += G) g->sstart[q] = s.sstart[q];
        if (sl == 0) { g->nic = nic; g->nsteps = s.nsteps; }
    }
    __device__ void restore_static(const S *g) {
        if (!g) { if (sl == 0) { s.nic = 0; s.nsteps = 0; } __syncwarp(); return; }       // idle lane group of a packed warp
        const int NPn = n * (n - 1) / 2, nic = g->nic, ns = g->nsteps;
        for (int q = sl; q < NPn; q += G) s.pair[q] = g->pair[q];
        for (int c = sl; c < nic; c += G) { s.ce0[c] = g->ce0[c]; s.ce1[c] = g->ce1[c]; s.ccode[c] = g->ccode[c]; s.cpos[c] = g->cpos[c]; }
        for (int q = sl; q <= ns; q += G) s.sstart[q] = g->sstart[q];
        if (sl == 0) { s.nic = nic; s.nsteps = ns; }
        __syncwarp();
    }
    __device__ void save_pairs(S *g) const {
        const int NPn = n * (n - 1) / 2;
        for (int q = sl; q < NPn; q += G) g->pair[q].w = s.pair[q].w;
    }

    // coupled path resume: the contact list was just re-detected (same positions => same list); fetch its impulses
    __device__ void load_contact_state(const float *ws_bounce, const float *ws_jbias) {
        for (int c = sl; c < s.nic; c += G) {
            size_t g = (size_t)o * SAPIO_ICAP + c;
            const int pos = s.cpos[c];
            s.cacc[pos] = make_float4(w.ic_jn[g], w.ic_jt[g], ws_jbias[g], 0.f); s.ce0[pos].w = ws_bounce[g];
        }
        __syncwarp();
    }
};

// ------------------------------------------------------------------------------------------------------------------
// organisms with no cross contact: the whole solve in one launch, everything on chip
// ------------------------------------------------------------------------------------------------------------------

#ifndef ORG_GC
#define ORG_GC 0          /* measured (profiles/r1c_org_solve.md): 15 instead of 8 warps per SM, same time -- kept as a switch */
#endif
#define ORG_GC_WARPS (148 * 32)             /* scratch slots per GC class: more than any grid of resident warps */
template <int CM> struct OrgCfg {
    static constexpr int G = CM <= 8 ? 4 : CM <= 16 ? 8 : CM <= 32 ? 16 : 32;          // lanes per organism
    static constexpr int NG = 32 / G;                                                     // organisms per warp
    static constexpr bool GC = ORG_GC && CM > 32;                                         // pin constants in global scratch
    static constexpr int WARP_BYTES = (int)sizeof(OrgS<CM, GC>) * NG;
    // an SM has 228 KB of shared memory, a CTA at most 227 KB, and every resident CTA reserves 1 KB: one CTA of up to 8 warps for the
    // large classes, two (or more) CTAs of up to 8 warps for the small ones
    static constexpr int FIT1 = (227 * 1024) / WARP_BYTES;                                // warps of a CTA that has the SM to itself
    static constexpr int FIT2 = ((228 * 1024 - 2 * 1024) / 2) / WARP_BYTES;               // warps per CTA when two CTAs share the SM
    static constexpr int CTAS = FIT1 <= 8 ? 1 : 2;                                        // CTAs per SM the kernel is compiled for (register budget)
    static constexpr int WARPS = FIT1 <= 8 ? (FIT1 < 1 ? 1 : FIT1) : (FIT2 < 8 ? FIT2 : 8);
};


template <int CM>
__global__ void __launch_bounds__(OrgCfg<CM>::WARPS * 32, OrgCfg<CM>::CTAS) k_org_solve(WORLD_ARG, Workspace W, StepK K, int cls) {
    constexpr int G = OrgCfg<CM>::G, NG = OrgCfg<CM>::NG;
    constexpr bool GC = OrgCfg<CM>::GC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = lane_id();
    OrgS<CM, GC> &s = reinterpret_cast<OrgS<CM, GC> *>(smem_raw)[(threadIdx.x >> 5) * NG + lane / G];
    float4 *cst = GC ? W.pair_cst + ((size_t)(cls - 4) * ORG_GC_WARPS + (size_t)blockIdx.x * OrgCfg<CM>::WARPS + (threadIdx.x >> 5)) * SAPIO_NPAIR : nullptr;
    const int start = W.flags[FL_CLS0 + cls], end = W.flags[FL_CLS0 + cls + 1];
    const int n_items = (end - start + NG - 1) / NG;
    const float dt_coef = w.g_stepped[0] ? 1.0f : 0.0f;   // prev_dt == 0 on the very first step
    const int iterations = w.p.iterations;
    int nic_sum = 0;
#ifdef ORG_PHASE_PROF
    long long ph[10] = {0}; long long tk = clock64(); int cnt_org = 0, sum_n = 0, sum_nic = 0, sum_ns = 0;
#define PH(k) do { long long t_ = clock64(); ph[k] += t_ - tk; tk = t_; } while (0)
#else
#define PH(k) do { } while (0)
#endif
    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(&W.flags[FL_WORK0 + cls], 1);          // dynamic scheduling: organisms differ in cost
        q = __shfl_sync(FULL, q, 0);
        if (q >= n_items) break;
        const int idx = end - 1 - (q * NG + lane / G);                        // largest organisms of the class first
        OrgCtx<CM, G, GC> c(s, w, K, idx >= start ? W.sorted_orgs[idx] : -1, cst);
        PH(0);
        c.load(); PH(1);
        c.prestep_pairs(); PH(2);
        c.schedule(); PH(3);
        c.template prestep_contacts<true>(); PH(4);
        const int nsteps_max = OrgCtx<CM, G, GC>::wmax(s.nsteps), tmax = 2 * OrgCtx<CM, G, GC>::wmax(c.n) - 3;
        if (dt_coef != 0.f) c.warm_start(dt_coef, nullptr);
        else c.zero_pin_acc();
        PH(5);
        for (int it = 0; it < iterations; it++) { c.template contacts_pass<false>(0.f, nsteps_max); PH(6); c.template joints_pass<false>(0.f, tmax); PH(7); }
        if (c.o >= 0) { c.store_vel(); c.store_contacts(nullptr, nullptr); if (c.sl == 0) nic_sum += s.nic; }
        __syncwarp();
        PH(8);
#ifdef ORG_PHASE_PROF
        cnt_org++; sum_n += c.n; sum_nic += s.nic; sum_ns += s.nsteps;
#endif
    }
#ifdef ORG_PHASE_PROF
    if (blockIdx.x == 0 && threadIdx.x == 0 && (w.g_tick[0] % 1000) == 7)
        printf("ORGPH cls %d tick %d orgs %d n %d nic %d nsteps %d | fetch %lld load %lld pairs %lld sched %lld cprestep %lld warm %lld contacts %lld joints %lld store %lld\n",
               cls, (int)w.g_tick[0], cnt_org, sum_n, sum_nic, sum_ns, ph[0], ph[1], ph[2], ph[3], ph[4], ph[5], ph[6], ph[7], ph[8]);
#endif
#undef PH
    if (nic_sum) atomicAdd(&W.flags[FL_NIC], nic_sum);
}

