// train.cuh -- neural.train_network (neural.py:369-423): the "dopamine RL update". One CTA per organism whose training
// is due. Layout of the work:
//   hash     : every short-term row and every stored row gets a 128-bit hash of its 1-decimal key vector (warp per row)
//   curate   : one warp walks the short-term rows chronologically on hashes, rewards and row indices only (no data moves)
//   apply    : the surviving row moves are materialised in parallel (sources and destinations are disjoint, see below)
//   epoch    : batch forward / backward with activations stored transposed [unit][sample] in the CTA's scratch (sample index
//              fastest: coalesced), one thread per (unit, sample); then one thread per optimizer parameter
//   curation : for every short-term memory (chronological), the best-rewarded action among short-term memories with the same
//              1-decimal input key replaces / joins the curated memory (first match by key; replaced when its reward is not
//              greater); then trim to memory_limit dropping the lowest reward first. Keys are Python round(x, 1) semantics on
//              the doubles the reference holds (fp32 values for short-term inputs, round(x, 3) values for stored inputs).
//              Storage policy as in oracle/brain.c: replace in place, trim moves the last row into the hole.
//   epoch    : Trainer.train_epoch (networks.py:75-87): full batch forward, MSELoss(mean), backward, torch.optim.Adam step on
//              the input and output layers only (hidden layers are not optimizer parameters, networks.py:221).
//              fp32 FMA, no tensor cores: batches are <= 200 x <= in_cap x <= width_cap and ragged per organism, and the
//              parity contract is 1e-5 against torch's fp32 (tf32/bf16 inputs cannot meet it).
#pragma once
#include "common.cuh"
#include "rules.cuh"

#ifndef TRAIN_THREADS
#define TRAIN_THREADS 256
#endif
#define TRAIN_CTAS_PER_SM (1024 / TRAIN_THREADS)
#define TRAIN_MAXW 128

// Python round(v, 1) * 10 for a double v, exactly (ties resolved on the exact binary value, half to even when exact)
__device__ __forceinline__ int py_round1_x10(double v) {
    double t = v * 10.0;
    double e = fma(v, 10.0, -t);                  // exact residual of the product
    double fl = floor(t);
    if (t - fl == 0.5) {
        if (e > 0.0) return (int)fl + 1;
        if (e < 0.0) return (int)fl;
        long long m = (long long)fl;
        return (int)((m & 1) ? m + 1 : m);        // exact tie: half to even
    }
    return (int)rint(t);
}
__device__ __forceinline__ int key_stm(float x) { return py_round1_x10((double)x); }                 // float(tensor element)
__device__ __forceinline__ int key_mem(float x) { return py_round1_x10(rint((double)x * 1000.0) / 1000.0); }   // round(x, 3) first

struct TrainWs { float *buf; size_t per_cta; };   // per CTA: (in_cap + 4 * TRAIN_MAXW + 4) * MEMROWS floats
#define TR_S SAPIO_MEMROWS                          // column stride of the transposed activation buffers

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {          // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
// key-vector hash: sum over positions of two independent mixes of (key, position). Rows are equal iff their key vectors are equal,
// up to a 2^-128 collision probability per comparison.
__device__ __forceinline__ void key_hash_add(int key, int c, unsigned long long &a, unsigned long long &b) {
    unsigned long long v = ((unsigned long long)(unsigned)c << 32) | (unsigned)(key & 0xffff);
    a += mix64(v + 0x9E3779B97F4A7C15ull); b += mix64(v * 0xD6E8FEB86659FD93ull + 0x2545F4914F6CDD1Dull);
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
    for (int s = 16; s; s >>= 1) v += __shfl_xor_sync(FULL, v, s);
    return v;
}

// Trainer.train_rnn (networks.py:56-73) as neural.train_network calls it (neural.py:407-413); see oracle/brain.c oracle_train_rnn for what
// the reference's loop amounts to: ONE sample -- input = the last row of the memory list, hidden = trainingHidden[M - 1] (the (M-1)-th entry
// ever appended), target = the first row's action -- MSE over the 4 outputs, plain SGD p += -lr * grad on the input and output layers, and
// network.lastHidden = the hidden head of that sample under the old weights. Whole CTA, one thread per output unit; scratch = the CTA's
// activation buffer (x | h0 | two ping-pong rows | hl copy | two gradient rows), seq_s = append stamps of the rows, red = a small shared array.
__device__ __forceinline__ void train_rnn_step(const SapioWorld &w, int o, int M, const int *seq_s, float *red, float *scratch) {
    const SapioParams &p = w.p;
    const int IN = p.in_cap, RC = p.rnn_cap, H = w.org_rnn_h[o], tid = threadIdx.x;
    const int32_t *ldim = w.org_ldim + o * 16;
    const int nl = w.org_nlayers[o], I = ldim[0], F0 = I + H, h0 = ldim[1], hl = ldim[nl - 1];
    float *x = scratch, *a0 = x + 256, *pa = a0 + TRAIN_MAXW, *pb = pa + TRAIN_MAXW, *ah = pb + TRAIN_MAXW, *da = ah + TRAIN_MAXW, *db = da + TRAIN_MAXW;
    float *brain = w.org_brain + (size_t)o * p.brain_cap;
    __shared__ int sh_last, sh_first;
    if (tid == 0) {
        int last = 0, first = 0;
        for (int m = 1; m < M; m++) { if (seq_s[m] > seq_s[last]) last = m; if (seq_s[m] < seq_s[first]) first = m; }
        sh_last = last; sh_first = first;
    }
    __syncthreads();
    const int last = sh_last, first = sh_first;
    for (int c = tid; c < F0; c += TRAIN_THREADS)
        x[c] = c < I ? w.org_mem_in[((size_t)o * SAPIO_MEMROWS + last) * IN + c] : (M - 1 < w.org_th_n[o]) ? w.org_th[((size_t)o * SAPIO_MEMROWS + (M - 1)) * RC + (c - I)] : 0.f;
    __syncthreads();
    // forward: layers 0 .. nl-1, then the hidden head on the last hidden activation
    const float *cur = x; float *dst = a0;
    int off = 0, off_out = 0;
    for (int l = 0; l < nl - 1; l++) {
        const int fin = l == 0 ? F0 : ldim[l], fout = ldim[l + 1];
        const float *Wm = brain + off, *b = Wm + fout * fin;
        for (int r = tid; r < fout; r += TRAIN_THREADS) {
            float s = 0.f;
            for (int c = 0; c < fin; c++) s += Wm[r * fin + c] * cur[c];
            s += b[r];
            if (l == 0) s = 1.0f / (1.0f + expf(-s));
            dst[r] = s;
        }
        __syncthreads();
        cur = dst; dst = (l == 0) ? pa : (dst == pa ? pb : pa);
        off += fout * fin + fout;
    }
    off_out = off;
    float *Wout = brain + off_out, *bout = Wout + 4 * hl, *Wh = bout + 4, *bh = Wh + H * hl;
    for (int c = tid; c < hl; c += TRAIN_THREADS) ah[c] = cur[c];                      // the last hidden activation (== a0 when there is one hidden layer)
    if (tid < 4 + H) {
        const float *Wm = tid < 4 ? Wout + tid * hl : Wh + (tid - 4) * hl;
        float s = 0.f;
        for (int c = 0; c < hl; c++) s += Wm[c] * cur[c];
        s += tid < 4 ? bout[tid] : bh[tid - 4];
        red[tid] = s;                                                                  // out[0..4), hid[0..H)
    }
    __syncthreads();
    if (tid < 4) {
        const float e = red[tid] - w.org_mem_out[((size_t)o * SAPIO_MEMROWS + first) * 4 + tid];
        red[32 + tid] = e * e; red[64 + tid] = 2.0f * e / 4.0f;                        // dO
    }
    if (tid >= 32 && tid < 32 + H) w.org_last_hidden[(size_t)o * RC + (tid - 32)] = red[4 + tid - 32];
    __syncthreads();
    if (tid == 0) w.org_loss[o] = (float)(((double)red[32] + (double)red[33] + (double)red[34] + (double)red[35]) / 4);
    const float dO0 = red[64], dO1 = red[65], dO2 = red[66], dO3 = red[67];
    // backward through the activation-free hidden chain to the first layer's output
    float *d = da, *d2 = db;
    for (int c = tid; c < hl; c += TRAIN_THREADS) { float s = 0.f; s += dO0 * Wout[0 * hl + c]; s += dO1 * Wout[1 * hl + c]; s += dO2 * Wout[2 * hl + c]; s += dO3 * Wout[3 * hl + c]; d[c] = s; }
    __syncthreads();
    for (int l = nl - 2; l >= 1; l--) {
        const int fin = ldim[l], fout = ldim[l + 1];
        int offl = 0;
        for (int t = 0; t < l; t++) offl += ldim[t + 1] * (t == 0 ? F0 : ldim[t]) + ldim[t + 1];
        const float *Wm = brain + offl;
        for (int c = tid; c < fin; c += TRAIN_THREADS) { float s = 0.f; for (int r = 0; r < fout; r++) s += d[r] * Wm[r * fin + c]; d2[c] = s; }
        __syncthreads();
        float *t2 = d; d = d2; d2 = t2;
    }
    // SGD: p = p + (-lr) * grad, products rounded like torch's (grad materialised in fp32 first)
    const float nlr = -p.learning_rate;
    float *Win = brain, *bin = brain + h0 * F0;
    for (int e = tid; e < 4 * hl + 4; e += TRAIN_THREADS) {
        if (e < 4 * hl) { const int q = e / hl, c = e - q * hl; const float g = __fmul_rn(red[64 + q], ah[c]); Wout[e] = __fadd_rn(Wout[e], __fmul_rn(nlr, g)); }
        else bout[e - 4 * hl] = __fadd_rn(bout[e - 4 * hl], __fmul_rn(nlr, red[64 + e - 4 * hl]));
    }
    for (int e = tid; e < h0 * F0 + h0; e += TRAIN_THREADS) {
        const int r = e < h0 * F0 ? e / F0 : e - h0 * F0;
        const float hv = a0[r], dz = __fmul_rn(__fmul_rn(d[r], hv), 1.0f - hv);
        if (e < h0 * F0) { const float g = __fmul_rn(dz, x[e - r * F0]); Win[e] = __fadd_rn(Win[e], __fmul_rn(nlr, g)); }
        else bin[r] = __fadd_rn(bin[r], __fmul_rn(nlr, dz));
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TRAIN_THREADS, TRAIN_CTAS_PER_SM) k_train(WORLD_ARG, RulesWs R, TrainWs TW) {
    __shared__ unsigned long long hs[2][SAPIO_STM], hm[2][SAPIO_STM], hmem[2][SAPIO_MEMROWS];
    __shared__ float rew_s[SAPIO_MEMROWS], red_f[TRAIN_THREADS];
    __shared__ int16_t rowsrc[SAPIO_MEMROWS];      // -1 row untouched | 0..231 content of original row | 256 + i short-term row i
    __shared__ uint8_t bestof[SAPIO_STM];
    __shared__ int seq_s[SAPIO_MEMROWS];           // list position of every curated row (append stamps): Trainer.train_rnn reads the last row and the first
    __shared__ int sh_M;
    const SapioParams &p = w.p;
    const int IN = p.in_cap, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float *XT = TW.buf + (size_t)blockIdx.x * TW.per_cta, *H0T = XT + (size_t)IN * TR_S, *P = H0T + TRAIN_MAXW * TR_S, *Q = P + TRAIN_MAXW * TR_S,
          *D = Q + TRAIN_MAXW * TR_S, *dOT = D + TRAIN_MAXW * TR_S;
    const int n_train = R.counters[RC_NTRAIN];
    __shared__ int sh_q;
    for (;;) {                                             // organisms differ by the size of their memory and of their brain: pulled, not dealt
        __syncthreads();
        if (tid == 0) sh_q = atomicAdd(&R.counters[RC_TRAIN_TICKET], 1);
        __syncthreads();
        const int q = sh_q;
        if (q >= n_train) break;
        if (tid == 0 && q >= (int)gridDim.x) atomicOr((unsigned long long *)&w.g_stats[SAPIO_STAT_RARE_PATHS], (unsigned long long)SAPIO_RARE_TRAIN_QUEUE);   // a CTA came back for more
        const int o = R.train_list[q];
        const int32_t *ldim = w.org_ldim + o * 16;
        const int nl = w.org_nlayers[o], I = ldim[0];
        const int sc = w.org_stm_n[o], sn = min(sc, SAPIO_STM), M0 = w.org_mem_n[o];
        float *min_ = w.org_mem_in + (size_t)o * SAPIO_MEMROWS * IN, *mout = w.org_mem_out + (size_t)o * SAPIO_MEMROWS * 4, *mrew = w.org_mem_rew + (size_t)o * SAPIO_MEMROWS;
        const float *sin_ = w.org_stm_in + (size_t)o * SAPIO_STM * IN, *sout = w.org_stm_out + o * SAPIO_STM * 4, *srew = w.org_stm_rew + o * SAPIO_STM;
#define STM_SLOT(i) ((sc - sn + (i)) % SAPIO_STM)
        // ---- hash: short-term rows under both key rules (as a probe: float(x); as a stored row: round(x, 3)), stored rows ----
        for (int r = wid; r < sn + M0; r += TRAIN_THREADS / 32) {
            unsigned long long a = 0, b = 0, a2 = 0, b2 = 0;
            if (r < sn) {
                const float *x = sin_ + (size_t)STM_SLOT(r) * IN;
                for (int c = lane; c < I; c += 32) { float v = x[c]; key_hash_add(key_stm(v), c, a, b); key_hash_add(key_mem(v), c, a2, b2); }
                a = warp_sum_u64(a); b = warp_sum_u64(b); a2 = warp_sum_u64(a2); b2 = warp_sum_u64(b2);
                if (lane == 0) { hs[0][r] = a; hs[1][r] = b; hm[0][r] = a2; hm[1][r] = b2; }
            } else {
                const int m = r - sn;
                const float *x = min_ + (size_t)m * IN;
                for (int c = lane; c < I; c += 32) key_hash_add(key_mem(x[c]), c, a, b);
                a = warp_sum_u64(a); b = warp_sum_u64(b);
                if (lane == 0) { hmem[0][m] = a; hmem[1][m] = b; }
            }
        }
        const int H = w.org_rnn_h[o], RC = p.rnn_cap;
        int32_t *mseq = w.org_mem_seq + (size_t)o * SAPIO_MEMROWS;
        for (int m = tid; m < SAPIO_MEMROWS; m += TRAIN_THREADS) { rowsrc[m] = -1; rew_s[m] = m < M0 ? mrew[m] : 0.f; seq_s[m] = m < M0 ? mseq[m] : 0; }
        __syncthreads();
        // (a) sorted(allActions, key=reward)[-1] among the short-term rows with the same key: highest reward, the LAST of equal maxima
        if (tid < sn) {
            float br = -INFINITY; int bj = tid;
            for (int j = 0; j < sn; j++)
                if (hs[0][j] == hs[0][tid] && hs[1][j] == hs[1][tid]) { float r = srew[STM_SLOT(j)]; if (r >= br) { br = r; bj = j; } }
            bestof[tid] = (uint8_t)bj;
        }
        __syncthreads();
        // ---- curate (neural.py:374-404): one warp, chronological, bookkeeping only ----
        if (wid == 0) {
            int M = M0, stamp = w.org_mem_stamp[o], thn = H > 0 ? w.org_th_n[o] : 0;
            const int hpop = H > 0 ? w.org_hm_n[2 * o + 1] : 0;
            for (int i = 0; i < sn; i++) {
                const unsigned long long a = hs[0][i], b = hs[1][i];
                int found = -1;                                   // (b) rounded_training_data.index(key): first stored row with the key
                for (int base = 0; base < M; base += 32) {
                    int m = base + lane;
                    unsigned bal = __ballot_sync(FULL, m < M && hmem[0][m] == a && hmem[1][m] == b);
                    if (bal) { found = base + __ffs(bal) - 1; break; }
                }
                const float best_r = srew[STM_SLOT(bestof[i])];
                int add = 1, dst = M;
                if (found >= 0) { if (rew_s[found] > best_r) add = 0; else dst = found; }
                if (add && dst == M) { if (M >= SAPIO_MEMROWS) { add = 0; if (lane == 0) stat_add(w, SAPIO_STAT_OVERFLOW, 1); } else M++; }   // table full (only without finite_memory): dropped and counted
                __syncwarp();
                if (add && lane == 0) {                           // input of THIS memory, action of the best one, reward of this one (sic, neural.py:395)
                    rowsrc[dst] = (int16_t)(256 + i); rew_s[dst] = srew[STM_SLOT(i)]; hmem[0][dst] = hm[0][i]; hmem[1][dst] = hm[1][i];
                    seq_s[dst] = stamp;                           // remove_memory + append: the row is now the last of the list
                }
                if (add) {
                    stamp++;
                    if (H > 0 && thn < SAPIO_MEMROWS) {           // trainingHidden.append(hidden_memory[i]) (neural.py:396-397): append-only, the first 232 entries kept
                        if (lane < H) w.org_th[((size_t)o * SAPIO_MEMROWS + thn) * RC + lane] = w.org_hm[((size_t)o * SAPIO_HM + (size_t)((hpop + i) % SAPIO_HM)) * RC + lane];
                        thn++;
                    }
                }
                __syncwarp();
            }
            // trim (neural.py:399-404): drop the lowest reward (first of equal minima), the last row fills the hole
            while (M > p.memory_limit && (w.org_flags[o] & SAPIO_F_FINITE_MEMORY)) {
                float lr = INFINITY; int li = 1 << 30;
                for (int m = lane; m < M; m += 32) { float r = rew_s[m]; if (r < lr) { lr = r; li = m; } }
                for (int s = 16; s; s >>= 1) {
                    float r2 = __shfl_xor_sync(FULL, lr, s); int j2 = __shfl_xor_sync(FULL, li, s);
                    if (r2 < lr || (r2 == lr && j2 < li)) { lr = r2; li = j2; }
                }
                const int last = M - 1, lo = li < M ? li : last;
                __syncwarp();
                if (lo != last && lane == 0) { rowsrc[lo] = rowsrc[last] < 0 ? (int16_t)last : rowsrc[last]; rew_s[lo] = rew_s[last]; seq_s[lo] = seq_s[last]; }
                M = last;
                __syncwarp();
            }
            if (lane == 0) { sh_M = M; w.org_mem_stamp[o] = stamp; if (H > 0) w.org_th_n[o] = thn; }
        }
        __syncthreads();
        const int M = sh_M;
        for (int m = tid; m < M; m += TRAIN_THREADS) mseq[m] = seq_s[m];
        // ---- apply: a row only ever moves when it is the last one, so every original source row lies at or beyond M and every
        //      destination below M: the copies are independent ----
        for (int m = wid; m < M; m += TRAIN_THREADS / 32) {
            const int src = rowsrc[m];
            if (src < 0) continue;
            const float *x, *y;
            if (src >= 256) { x = sin_ + (size_t)STM_SLOT(src - 256) * IN; y = sout + STM_SLOT(bestof[src - 256]) * 4; }
            else { x = min_ + (size_t)src * IN; y = mout + src * 4; }
            for (int c = lane; c < I; c += 32) min_[(size_t)m * IN + c] = x[c];
            if (lane < 4) mout[m * 4 + lane] = y[lane];
            if (lane == 0) mrew[m] = rew_s[m];
        }
        if (tid == 0) w.org_mem_n[o] = M;
        __syncthreads();
        if (M > 0 && H > 0) train_rnn_step(w, o, M, seq_s, red_f, XT);
        else if (M > 0) {
            float *brain = w.org_brain + (size_t)o * p.brain_cap;
            const int h0 = ldim[1], hl = ldim[nl - 1];
            int off_out = 0;
            for (int l = 0; l < nl - 1; l++) off_out += ldim[l + 1] * ldim[l] + ldim[l + 1];
            float *Win = brain, *bin = brain + h0 * I, *Wout = brain + off_out, *bout = Wout + 4 * hl;
            // ---- epoch, forward: X^T, then layer by layer; thread = (unit r, sample m), m fastest ----
            const int M4 = (M + 3) >> 2;                          // sample groups of four (the padding samples are zeros and never read back)
            for (int e = tid; e < 4 * M4 * I; e += TRAIN_THREADS) { int m = e / I, c = e - m * I; XT[c * TR_S + m] = m < M ? min_[(size_t)m * IN + c] : 0.f; }
            __syncthreads();
            const float *cur = XT, *hlastT = XT;
            int off = 0;
            for (int l = 0; l < nl; l++) {
                const int fin = ldim[l], fout = ldim[l + 1];
                const float *Wm = brain + off, *b = Wm + fout * fin;
                float *dst = (l == nl - 1) ? dOT : (l == 0) ? H0T : (cur == P ? Q : P);
                // four samples per thread: one weight load feeds four independent FMA chains, activations move as float4
                // (columns are padded to a multiple of 4 samples; per output the summation order over c is unchanged)
                for (int e = tid; e < fout * M4; e += TRAIN_THREADS) {
                    const int r = e / M4, m = 4 * (e - r * M4);
                    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int c = 0; c < fin; c++) {
                        const float wv = Wm[r * fin + c];
                        const float4 a = *reinterpret_cast<const float4 *>(cur + c * TR_S + m);
                        s.x += wv * a.x; s.y += wv * a.y; s.z += wv * a.z; s.w += wv * a.w;
                    }
                    const float bv = b[r];
                    s.x += bv; s.y += bv; s.z += bv; s.w += bv;
                    if (l == 0) { s.x = 1.0f / (1.0f + expf(-s.x)); s.y = 1.0f / (1.0f + expf(-s.y)); s.z = 1.0f / (1.0f + expf(-s.z)); s.w = 1.0f / (1.0f + expf(-s.w)); }
                    *reinterpret_cast<float4 *>(dst + r * TR_S + m) = s;
                }
                __syncthreads();
                if (l == nl - 2) hlastT = dst;
                cur = dst; off += fout * fin + fout;
            }
            // MSELoss(mean) and its gradient at the outputs (thread per sample)
            float lsum = 0.f;
            for (int m = tid; m < M; m += TRAIN_THREADS) {
                const float *y = mout + m * 4;
#pragma unroll
                for (int z = 0; z < 4; z++) { float e = dOT[z * TR_S + m] - y[z]; lsum += e * e; dOT[z * TR_S + m] = 2.0f * e / (float)(M * 4); }
            }
            red_f[tid] = lsum;
            __syncthreads();
            for (int s = TRAIN_THREADS / 2; s; s >>= 1) { if (tid < s) red_f[tid] += red_f[tid + s]; __syncthreads(); }
            if (tid == 0) w.org_loss[o] = red_f[0] / (float)(M * 4);
            // ---- backward down to the first layer's pre-activation (hidden layers have no activation) ----
            float *d = D, *d2 = (hlastT == P) ? Q : P;
            for (int e = tid; e < hl * M4; e += TRAIN_THREADS) {
                const int c = e / M4, m = 4 * (e - c * M4);
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int z = 0; z < 4; z++) {
                    const float wv = Wout[z * hl + c];
                    const float4 a = *reinterpret_cast<const float4 *>(dOT + z * TR_S + m);
                    s.x += a.x * wv; s.y += a.y * wv; s.z += a.z * wv; s.w += a.w * wv;
                }
                *reinterpret_cast<float4 *>(d + c * TR_S + m) = s;
            }
            __syncthreads();
            for (int l = nl - 2; l >= 1; l--) {
                const int fin = ldim[l], fout = ldim[l + 1];
                int offl = 0;
                for (int t = 0; t < l; t++) offl += ldim[t + 1] * ldim[t] + ldim[t + 1];
                const float *Wm = brain + offl;
                for (int e = tid; e < fin * M4; e += TRAIN_THREADS) {
                    const int c = e / M4, m = 4 * (e - c * M4);
                    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int r = 0; r < fout; r++) {
                        const float wv = Wm[r * fin + c];
                        const float4 a = *reinterpret_cast<const float4 *>(d + r * TR_S + m);
                        s.x += a.x * wv; s.y += a.y * wv; s.z += a.z * wv; s.w += a.w * wv;
                    }
                    *reinterpret_cast<float4 *>(d2 + c * TR_S + m) = s;
                }
                __syncthreads();
                float *t2 = d; d = d2; d2 = t2;
            }
            float *dz0T = d2;
            for (int e = tid; e < h0 * 4 * M4; e += TRAIN_THREADS) { const int r = e / (4 * M4), m = e - r * 4 * M4; float hv = H0T[r * TR_S + m]; dz0T[r * TR_S + m] = d[r * TR_S + m] * hv * (1.0f - hv); }
            __syncthreads();
            // ---- one thread per optimizer parameter: gradient (samples in ascending order) and the Adam step ----
            const int t = w.org_adam_t[o] + 1;
            const float lr = p.learning_rate, b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
            const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
            const float step_size = (float)((double)lr / bc1), bc2s = (float)sqrt(bc2);
            float *am = w.org_adam_m + (size_t)o * p.adam_cap, *av = w.org_adam_v + (size_t)o * p.adam_cap;
            const int np_in = h0 * I + h0, np_out = 4 * hl + 4;
            for (int e = tid; e < np_in + np_out; e += TRAIN_THREADS) {
                // Adam divides by |g| + 1e-8: gradients are summed over the batch in double so that their own rounding noise does not
                // decide the update of near-zero components
                // both factors are stored sample-fastest (dz0T / dOT rows, X^T and the last hidden layer's activations): four samples per
                // pair of 16-byte loads, added in ascending sample order like the scalar loop (the padding samples beyond M are skipped)
                double gd = 0.0; float *prm;
                const float *fa, *fb = nullptr;
                if (e < h0 * I) { const int r = e / I, c = e % I; fa = dz0T + r * TR_S; fb = XT + c * TR_S; prm = &Win[e]; }
                else if (e < np_in) { const int r = e - h0 * I; fa = dz0T + r * TR_S; prm = &bin[r]; }
                else if (e < np_in + 4 * hl) { const int z = (e - np_in) / hl, c = (e - np_in) % hl; fa = dOT + z * TR_S; fb = hlastT + c * TR_S; prm = &Wout[e - np_in]; }
                else { const int z = e - np_in - 4 * hl; fa = dOT + z * TR_S; prm = &bout[z]; }
                const int Mq = M & ~3;
                if (fb) {
                    for (int m = 0; m < Mq; m += 4) {
                        const float4 a = *reinterpret_cast<const float4 *>(fa + m), b = *reinterpret_cast<const float4 *>(fb + m);
                        gd += (double)__fmul_rn(a.x, b.x); gd += (double)__fmul_rn(a.y, b.y); gd += (double)__fmul_rn(a.z, b.z); gd += (double)__fmul_rn(a.w, b.w);
                    }
                    for (int m = Mq; m < M; m++) gd += (double)__fmul_rn(fa[m], fb[m]);
                } else {
                    for (int m = 0; m < Mq; m += 4) { const float4 a = *reinterpret_cast<const float4 *>(fa + m); gd += (double)a.x; gd += (double)a.y; gd += (double)a.z; gd += (double)a.w; }
                    for (int m = Mq; m < M; m++) gd += (double)fa[m];
                }
                const float g = (float)gd;
                float m1 = b1 * am[e] + (1.0f - b1) * g, v1 = b2 * av[e] + (1.0f - b2) * g * g;
                am[e] = m1; av[e] = v1;
                float denom = sqrtf(v1) / bc2s + eps;
                *prm = *prm - step_size * (m1 / denom);
            }
            __syncthreads();
            if (tid == 0) w.org_adam_t[o] = t;
        }
        if (tid == 0) { w.org_train_tick[o] = (int32_t)w.g_tick[0]; stat_add(w, SAPIO_STAT_TRAINS, 1); }
        __syncthreads();
#undef STM_SLOT
    }
}
