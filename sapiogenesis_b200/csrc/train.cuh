// train.cuh -- neural.train_network (neural.py:369-423): the "dopamine RL update". One CTA per organism whose training
// is due.
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

#define TRAIN_THREADS 256
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

struct TrainWs { float *dz0, *hl, *dO; };          // per CTA: [MEMROWS x TRAIN_MAXW], [MEMROWS x TRAIN_MAXW], [MEMROWS x 4]

__global__ void __launch_bounds__(TRAIN_THREADS) k_train(WORLD_ARG, RulesWs R, TrainWs TW) {
    extern __shared__ int16_t train_keys[];        // [STM][IN] short-term keys, then [MEMROWS][IN] memory keys
    __shared__ float red_f[TRAIN_THREADS];
    __shared__ int red_i[TRAIN_THREADS];
    __shared__ int sh_best, sh_found, sh_dst, sh_add, sh_M;
    const SapioParams &p = w.p;
    const int IN = p.in_cap, tid = threadIdx.x;
    int16_t *kstm = train_keys, *kmem = train_keys + SAPIO_STM * IN;
    float *dz0 = TW.dz0 + (size_t)blockIdx.x * SAPIO_MEMROWS * TRAIN_MAXW, *hlast = TW.hl + (size_t)blockIdx.x * SAPIO_MEMROWS * TRAIN_MAXW;
    float *dOut = TW.dO + (size_t)blockIdx.x * SAPIO_MEMROWS * 4;
    const int n_train = R.counters[RC_NTRAIN];
    for (int q = blockIdx.x; q < n_train; q += gridDim.x) {
        const int o = R.train_list[q];
        const int32_t *ldim = w.org_ldim + o * 16;
        const int nl = w.org_nlayers[o], I = ldim[0];
        const int sc = w.org_stm_n[o], sn = min(sc, SAPIO_STM);
        float *min_ = w.org_mem_in + (size_t)o * SAPIO_MEMROWS * IN, *mout = w.org_mem_out + (size_t)o * SAPIO_MEMROWS * 4, *mrew = w.org_mem_rew + (size_t)o * SAPIO_MEMROWS;
        const float *sin_ = w.org_stm_in + (size_t)o * SAPIO_STM * IN, *sout = w.org_stm_out + o * SAPIO_STM * 4, *srew = w.org_stm_rew + o * SAPIO_STM;
        if (tid == 0) sh_M = w.org_mem_n[o];
        for (int e = tid; e < sn * I; e += TRAIN_THREADS) { int r = e / I, c = e % I; int slot = (sc - sn + r) % SAPIO_STM; kstm[r * IN + c] = (int16_t)key_stm(sin_[(size_t)slot * IN + c]); }
        __syncthreads();
        for (int e = tid; e < sh_M * I; e += TRAIN_THREADS) { int r = e / I, c = e % I; kmem[r * IN + c] = (int16_t)key_mem(min_[(size_t)r * IN + c]); }
        __syncthreads();
        for (int i = 0; i < sn; i++) {                                             // chronological (neural.py:374)
            const int si = (sc - sn + i) % SAPIO_STM;
            // (a) sorted(allActions, key=reward)[-1]: highest reward, the LAST of equal maxima
            float br = -INFINITY; int bj = -1;
            if (tid < sn) {
                bool same = true;
                for (int c = 0; c < I && same; c++) same = kstm[tid * IN + c] == kstm[i * IN + c];
                if (same) { br = srew[(sc - sn + tid) % SAPIO_STM]; bj = tid; }
            }
            red_f[tid] = br; red_i[tid] = bj;
            __syncthreads();
            for (int s = TRAIN_THREADS / 2; s; s >>= 1) {
                if (tid < s) { float r2 = red_f[tid + s]; int j2 = red_i[tid + s]; if (j2 >= 0 && (red_i[tid] < 0 || r2 > red_f[tid] || (r2 == red_f[tid] && j2 > red_i[tid]))) { red_f[tid] = r2; red_i[tid] = j2; } }
                __syncthreads();
            }
            const float best_r = red_f[0]; const int best = (sc - sn + red_i[0]) % SAPIO_STM;
            __syncthreads();
            // (b) rounded_training_data.index(key): first stored row with the same key
            int found = 1 << 30;
            for (int m = tid; m < sh_M; m += TRAIN_THREADS) {
                bool same = true;
                for (int c = 0; c < I && same; c++) same = kmem[m * IN + c] == kstm[i * IN + c];
                if (same) { found = m; break; }
            }
            red_i[tid] = found;
            __syncthreads();
            for (int s = TRAIN_THREADS / 2; s; s >>= 1) { if (tid < s) red_i[tid] = min(red_i[tid], red_i[tid + s]); __syncthreads(); }
            if (tid == 0) {
                int f = red_i[0], add = 1, dst = sh_M;
                if (f < (1 << 30)) { if (mrew[f] > best_r) add = 0; else dst = f; }
                if (add && dst == sh_M) { if (sh_M >= SAPIO_MEMROWS) add = 0; else sh_M++; }
                sh_add = add; sh_dst = dst;
            }
            __syncthreads();
            if (sh_add) {
                const int dst = sh_dst;
                for (int c = tid; c < I; c += TRAIN_THREADS) { float v = sin_[(size_t)si * IN + c]; min_[(size_t)dst * IN + c] = v; kmem[dst * IN + c] = (int16_t)key_mem(v); }
                if (tid < 4) mout[dst * 4 + tid] = sout[best * 4 + tid];
                if (tid == 0) mrew[dst] = srew[si];                                                // sic: mem[2], not bestReward (neural.py:395)
            }
            __syncthreads();
        }
        // trim (neural.py:399-404)
        while (sh_M > p.memory_limit && (w.org_flags[o] & SAPIO_F_FINITE_MEMORY)) {
            const int M = sh_M;
            float lr = INFINITY; int li = 1 << 30;
            for (int m = tid; m < M; m += TRAIN_THREADS) { float r = mrew[m]; if (r < lr) { lr = r; li = m; } }
            red_f[tid] = lr; red_i[tid] = li;
            __syncthreads();
            for (int s = TRAIN_THREADS / 2; s; s >>= 1) {
                if (tid < s) { float r2 = red_f[tid + s]; int j2 = red_i[tid + s]; if (r2 < red_f[tid] || (r2 == red_f[tid] && j2 < red_i[tid])) { red_f[tid] = r2; red_i[tid] = j2; } }
                __syncthreads();
            }
            const int lo = red_i[0], last = M - 1;
            __syncthreads();
            if (lo != last) {
                for (int c = tid; c < I; c += TRAIN_THREADS) { min_[(size_t)lo * IN + c] = min_[(size_t)last * IN + c]; kmem[lo * IN + c] = kmem[last * IN + c]; }
                if (tid < 4) mout[lo * 4 + tid] = mout[last * 4 + tid];
                if (tid == 0) mrew[lo] = mrew[last];
            }
            if (tid == 0) sh_M = last;
            __syncthreads();
        }
        const int M = sh_M;
        if (tid == 0) w.org_mem_n[o] = M;
        if (M > 0) {
            float *brain = w.org_brain + (size_t)o * p.brain_cap;
            const int h0 = ldim[1], hl = ldim[nl - 1];
            int off_out = 0;
            for (int l = 0; l < nl - 1; l++) off_out += ldim[l + 1] * ldim[l] + ldim[l + 1];
            float *Win = brain, *bin = brain + h0 * I, *Wout = brain + off_out, *bout = Wout + 4 * hl;
            // phase 1: one thread per sample -- forward, loss, backward down to the first layer's pre-activation
            float lsum = 0.f;
            for (int m = tid; m < M; m += TRAIN_THREADS) {
                const float *x = min_ + (size_t)m * IN, *y = mout + m * 4;
                float a[2][TRAIN_MAXW];
                float *h0v = dz0 + (size_t)m * TRAIN_MAXW;           // holds H0 first, dz0 afterwards
                const float *cur = x; int off = 0, pp = 0;
                for (int l = 0; l < nl; l++) {
                    const int fin = ldim[l], fout = ldim[l + 1];
                    const float *Wm = brain + off, *b = Wm + fout * fin;
                    for (int r = 0; r < fout; r++) {
                        float s = 0.f;
                        for (int c = 0; c < fin; c++) s += Wm[r * fin + c] * cur[c];
                        s += b[r];
                        if (l == 0) { s = 1.0f / (1.0f + expf(-s)); h0v[r] = s; }
                        a[pp][r] = s;
                    }
                    if (l == nl - 2 || (nl == 1)) for (int r = 0; r < fout; r++) hlast[(size_t)m * TRAIN_MAXW + r] = a[pp][r];
                    cur = a[pp]; pp ^= 1; off += fout * fin + fout;
                }
                float dO[4];
#pragma unroll
                for (int z = 0; z < 4; z++) { float e = cur[z] - y[z]; lsum += e * e; dO[z] = 2.0f * e / (float)(M * 4); dOut[m * 4 + z] = dO[z]; }
                float *d = a[0], *d2 = a[1];
                for (int c = 0; c < hl; c++) { float s = 0.f;
#pragma unroll
                    for (int z = 0; z < 4; z++) s += dO[z] * Wout[z * hl + c];
                    d[c] = s; }
                for (int l = nl - 2; l >= 1; l--) {
                    const int fin = ldim[l], fout = ldim[l + 1];
                    int offl = 0;
                    for (int t = 0; t < l; t++) offl += ldim[t + 1] * ldim[t] + ldim[t + 1];
                    const float *Wm = brain + offl;
                    for (int c = 0; c < fin; c++) { float s = 0.f; for (int r = 0; r < fout; r++) s += d[r] * Wm[r * fin + c]; d2[c] = s; }
                    float *t2 = d; d = d2; d2 = t2;
                }
                for (int r = 0; r < h0; r++) { float hv = h0v[r]; h0v[r] = d[r] * hv * (1.0f - hv); }
            }
            red_f[tid] = lsum;
            __syncthreads();
            for (int s = TRAIN_THREADS / 2; s; s >>= 1) { if (tid < s) red_f[tid] += red_f[tid + s]; __syncthreads(); }
            if (tid == 0) w.org_loss[o] = red_f[0] / (float)(M * 4);
            __threadfence_block();
            __syncthreads();
            // phase 2: one thread per optimizer parameter -- gradient (samples in ascending order) and the Adam step
            const int t = w.org_adam_t[o] + 1;
            const float lr = p.learning_rate, b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
            const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
            const float step_size = (float)((double)lr / bc1), bc2s = (float)sqrt(bc2);
            float *am = w.org_adam_m + (size_t)o * p.adam_cap, *av = w.org_adam_v + (size_t)o * p.adam_cap;
            const int np_in = h0 * I + h0, np_out = 4 * hl + 4;
            for (int e = tid; e < np_in + np_out; e += TRAIN_THREADS) {
                // Adam divides by |g| + 1e-8: gradients are summed over the batch in double so that their own rounding noise does not
                // decide the update of near-zero components
                double gd = 0.0; float *prm;
                if (e < h0 * I) { int r = e / I, c = e % I; for (int m = 0; m < M; m++) gd += (double)__fmul_rn(dz0[(size_t)m * TRAIN_MAXW + r], min_[(size_t)m * IN + c]); prm = &Win[e]; }
                else if (e < np_in) { int r = e - h0 * I; for (int m = 0; m < M; m++) gd += (double)dz0[(size_t)m * TRAIN_MAXW + r]; prm = &bin[r]; }
                else if (e < np_in + 4 * hl) { int z = (e - np_in) / hl, c = (e - np_in) % hl; for (int m = 0; m < M; m++) gd += (double)__fmul_rn(dOut[m * 4 + z], hlast[(size_t)m * TRAIN_MAXW + c]); prm = &Wout[e - np_in]; }
                else { int z = e - np_in - 4 * hl; for (int m = 0; m < M; m++) gd += (double)dOut[m * 4 + z]; prm = &bout[z]; }
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
    }
}
