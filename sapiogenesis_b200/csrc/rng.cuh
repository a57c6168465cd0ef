// rng.cuh -- counter-based Philox4x32-10, the product's implementation of the RNG spec the oracle states in
// oracle/philox.h (same constants, same counter layout; defined by the spec, not shared by linkage).
// u = U(seed, uid, tick, purpose, idx): key = (seed_lo, seed_hi), counter = (uid_lo, uid_hi, tick, purpose << 24 | idx);
// a uniform is the top 24 bits of one output word times 2^-24 (exact in fp32).
#pragma once
#include <stdint.h>

enum {
    RNG_REPRO_GATE = 1,   // random.random()*100 <= health%            sprites.py:1665
    RNG_BOREDOM = 2,      // random.random() <= network.boredom        neural.py:294
    RNG_RANDN = 3,        // torch.randn(4)                            neural.py:295
    RNG_BEARINGS = 4,     // randomizeList(range(360))                 sprites.py:199-201
    RNG_GENOME = 5,       // DNA.randomize / DNA.mutate draw stream    sprites.py:872,1068
    RNG_BRAIN_INIT = 6,   // nn.Linear default init                    networks.py:220-233
    RNG_WEIGHT_FILL = 7,  // torch.rand rows of mutate_brain_layers    sprites.py:1041
    RNG_SYNTH = 8,        // synthetic world placement
    RNG_GENOME_INIT = 9,  // nn.Linear init of the tempNet in _setup_brain_structure  sprites.py:868
    RNG_WEIGHT_GARBAGE = 10, // uninitialised tail of resize_()d rows  sprites.py:1039
    RNG_EVENTS = 11,      // random dead cell event: gate, x, y (uid -1) Sapiogenesis.py:48-53
    RNG_MIGRATE = 12      // island migration: start of the emigrant walk (uid -2 - rank), placement candidates of an arrival (ours)
};

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__host__ __device__ __forceinline__ void rng4(uint64_t seed, int64_t uid, int32_t tick, uint32_t purpose, uint32_t idx, uint32_t out[4]) {
    out[0] = (uint32_t)(uint64_t)uid; out[1] = (uint32_t)((uint64_t)uid >> 32);
    out[2] = (uint32_t)tick; out[3] = (purpose << 24) | (idx & 0xFFFFFFu);
    philox4x32_10(out, (uint32_t)seed, (uint32_t)(seed >> 32));
}
__host__ __device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// sequential stream: draw n is word (n & 3) of block (n >> 2)
struct RngStream {
    uint64_t seed; int64_t uid; int32_t tick; uint32_t purpose; uint32_t n;
    uint32_t buf[4]; uint32_t have;   // block index cached in buf (+1), 0 = none
    __host__ __device__ RngStream(uint64_t s, int64_t u, int32_t t, uint32_t p, uint32_t start = 0) : seed(s), uid(u), tick(t), purpose(p), n(start), have(0) {}
    __host__ __device__ float next() {
        uint32_t blk = n >> 2;
        if (have != blk + 1) { rng4(seed, uid, tick, purpose, blk, buf); have = blk + 1; }
        float u = u01(buf[n & 3]);
        n++;
        return u;
    }
};
