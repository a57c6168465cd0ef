// grid.cuh -- the broadphase grid of Space.step and of the sensors: the living bodies counting-sorted into a dense uniform grid.
//
//   k_integrate          (bodies.cuh) integrates every body, writes its cell key, counts the cell (atomic) and appends the body to the
//                        list of living bodies (warp-aggregated, unordered)
//   scan_exclusive_i32   cell counts -> cell starts (three small kernels: tile sums, one CTA over the tile sums, tiles again)
//   k_grid_scatter       every living body takes a place in its cell's run (atomic countdown of the count: the counts are zero again
//                        afterwards)
//   k_grid_rank          within a run the bodies are put in ascending id -- the order a stable sort by cell of the slot array gives, so
//                        everything downstream (partner lists, contact keys, sensor views) is the same list bit for bit whatever the order
//                        of the atomics -- and the body's circle is written next to its id, so that the narrowphase reads runs of 16-byte
//                        records instead of gathering positions and sizes
//
// Work and traffic follow the ~3 M living bodies and the grid cells, not the 10 M body slots; no library sort or scan is involved.
#pragma once
#include "common.cuh"
#include "steptypes.cuh"

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ int scan_block_exclusive(int v, int &total) {      // exclusive scan of one value per thread over a 256-thread CTA
    __shared__ int wsum[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
    for (int s = 1; s < 32; s <<= 1) { int t = __shfl_up_sync(FULL, inc, s); if (lane >= s) inc += t; }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < SCAN_THREADS / 32; k++) { const int t = wsum[k]; if (k < wid) base += t; tot += t; }
    __syncthreads();
    total = tot;
    return base + inc - v;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const int *__restrict__ in, int n, int *__restrict__ parts) {
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int s = 0;
    if (base + SCAN_ITEMS <= n) {
        const int4 a = *reinterpret_cast<const int4 *>(in + base), b = *reinterpret_cast<const int4 *>(in + base + 4);
        s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    } else for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) s += in[base + k];
    int tot; scan_block_exclusive(s, tot);
    if (threadIdx.x == 0) parts[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_parts(int *__restrict__ parts, int ntiles) {      // one CTA: exclusive, in place
    int carry = 0;
    for (int base = 0; base < ntiles; base += SCAN_THREADS * 4) {                      // four consecutive tile sums per thread
        const int i = base + threadIdx.x * 4;
        int v[4], s = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { const int t = i + k < ntiles ? parts[i + k] : 0; v[k] = s; s += t; }
        int tot; const int ex = scan_block_exclusive(s, tot) + carry;
#pragma unroll
        for (int k = 0; k < 4; k++) if (i + k < ntiles) parts[i + k] = ex + v[k];
        carry += tot;
    }
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const int *__restrict__ in, int *__restrict__ out, int n, const int *__restrict__ parts) {
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    if (base + SCAN_ITEMS <= n) {
        const int4 a = *reinterpret_cast<const int4 *>(in + base), b = *reinterpret_cast<const int4 *>(in + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else for (int k = 0; k < SCAN_ITEMS; k++) v[k] = base + k < n ? in[base + k] : 0;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { const int t = v[k]; v[k] = s; s += t; }
    int tot; const int off = scan_block_exclusive(s, tot) + parts[blockIdx.x];
    if (base + SCAN_ITEMS <= n) {
        *reinterpret_cast<int4 *>(out + base) = make_int4(v[0] + off, v[1] + off, v[2] + off, v[3] + off);
        *reinterpret_cast<int4 *>(out + base + 4) = make_int4(v[4] + off, v[5] + off, v[6] + off, v[7] + off);
    } else for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) out[base + k] = v[k] + off;
}
// out[i] = in[0] + .. + in[i - 1], i < n (in and out 16-byte aligned, distinct; parts: (n + SCAN_TILE - 1) / SCAN_TILE ints)
static inline void scan_exclusive_i32(const int *in, int *out, int n, int *parts, cudaStream_t st) {
    const int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_tiles<<<ntiles, SCAN_THREADS, 0, st>>>(in, n, parts);
    k_scan_parts<<<1, SCAN_THREADS, 0, st>>>(parts, ntiles);
    k_scan_apply<<<ntiles, SCAN_THREADS, 0, st>>>(in, out, n, parts);
}

__global__ void __launch_bounds__(256) k_grid_scatter(Workspace W, const int *__restrict__ n_live) {
    const int n = *n_live;
    const int lane = threadIdx.x & 31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t b = W.gval[i], key = W.gkey[b];
        // neighbours in the list are mostly cells of one organism in one grid cell: one atomic per (warp, cell)
        const uint32_t peers = __match_any_sync(__activemask(), key);
        const int leader = __ffs(peers) - 1;
        int top = 0;
        if (lane == leader) top = atomicSub(&W.cell_cnt[key], __popc(peers));
        top = __shfl_sync(peers, top, leader);
        W.stmp[W.cell_start[key] + top - 1 - __popc(peers & ((1u << lane) - 1u))] = b;
    }
}
__global__ void __launch_bounds__(256) k_grid_rank(WORLD_ARG, Workspace W, const int *__restrict__ n_live) {
    const int n = *n_live, NC = w.p.n_org * MAXC;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t b = W.stmp[i], key = W.gkey[b];
        const int s = W.cell_start[key], e = W.cell_start[key + 1];
        int r = 0;
        for (int q = s; q < e; q++) r += W.stmp[q] < b;
        float2 p; float rad;
        if ((int)b < NC) { p = reinterpret_cast<const float2 *>(w.c_pos)[b]; rad = w.c_size[b] * 0.5f; }
        else { const int d = (int)b - NC; p = reinterpret_cast<const float2 *>(w.d_pos)[d]; rad = w.d_size[d] * 0.5f; }
        W.sval[s + r] = b; W.skey[s + r] = key;
        W.srec[s + r] = make_float4(p.x, p.y, rad, __int_as_float((int)b));
    }
}

// ---- deterministic reduction of doubles (sum or product): fixed assignment of elements to threads and a fixed tree, so that the world
//      stays bit-reproducible from its seed; two launches, the second one CTA over at most REDUCE_CTAS partial results -----------------
#define REDUCE_CTAS 296
template <bool MUL> __device__ __forceinline__ double reduce_op(double a, double b) { return MUL ? a * b : a + b; }
template <bool MUL>
__global__ void __launch_bounds__(256) k_reduce_d(const double *__restrict__ in, int n, double *__restrict__ out) {
    __shared__ double sh[256];
    double acc = MUL ? 1.0 : 0.0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) acc = reduce_op<MUL>(acc, in[i]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s; s >>= 1) { if (threadIdx.x < s) sh[threadIdx.x] = reduce_op<MUL>(sh[threadIdx.x], sh[threadIdx.x + s]); __syncthreads(); }
    if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}
// result in out[0]; parts: REDUCE_CTAS doubles of scratch
template <bool MUL>
static inline void reduce_doubles(const double *in, int n, double *out, double *parts, cudaStream_t st) {
    int ctas = (n + 255) / 256; if (ctas > REDUCE_CTAS) ctas = REDUCE_CTAS; if (ctas < 1) ctas = 1;
    k_reduce_d<MUL><<<ctas, 256, 0, st>>>(in, n, parts);
    k_reduce_d<MUL><<<1, 256, 0, st>>>(parts, ctas, out);
}
