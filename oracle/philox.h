/* philox.h -- counter-based RNG spec shared (by definition, not by linkage) between the oracle and the
 * CUDA kernels. TEST INFRASTRUCTURE: part of the CPU oracle; the product re-implements the same spec in
 * sapiogenesis_b200/csrc/rng.cuh.
 *
 * The reference draws from three unseeded global generators (Python `random`, numpy.random, torch;
 * SURVEY.md Appendix C-9). "Bit-exact decisions" are therefore only defined on identical uniforms:
 * every draw the rules make is u = U(seed, uid, tick, purpose, idx), a pure function of its coordinates.
 *
 * Philox4x32-10 (Salmon et al., SC'11): key = (seed_lo, seed_hi),
 * counter = (uid_lo, uid_hi, tick, purpose << 24 | idx). A uniform is the top 24 bits of one output
 * word times 2^-24, so float and double consumers see the same value exactly.
 */
#ifndef SAPIO_ORACLE_PHILOX_H
#define SAPIO_ORACLE_PHILOX_H
#include <stdint.h>

enum {
    SAPIO_RNG_REPRO_GATE = 1,   /* random.random()*100 <= health%            sprites.py:1665 */
    SAPIO_RNG_BOREDOM = 2,      /* random.random() <= network.boredom        neural.py:294   */
    SAPIO_RNG_RANDN = 3,        /* torch.randn(4)                            neural.py:295   */
    SAPIO_RNG_BEARINGS = 4,     /* randomizeList(range(360))                 sprites.py:199-201 */
    SAPIO_RNG_GENOME = 5,       /* DNA.randomize / DNA.mutate draw stream    sprites.py:872,1068 */
    SAPIO_RNG_BRAIN_INIT = 6,   /* nn.Linear default init                    networks.py:220-233 */
    SAPIO_RNG_WEIGHT_FILL = 7,  /* torch.rand rows of mutate_brain_layers    sprites.py:1041 */
    SAPIO_RNG_SYNTH = 8,        /* synthetic world placement (ours) */
    SAPIO_RNG_GENOME_INIT = 9,  /* nn.Linear init of the tempNet in DNA._setup_brain_structure  sprites.py:868 */
    SAPIO_RNG_WEIGHT_GARBAGE = 10, /* uninitialised tail of resize_()d rows in mutate_brain_layers sprites.py:1039 */
    SAPIO_RNG_EVENTS = 11,
    SAPIO_RNG_MIGRATE = 12  /* island migration: start of the emigrant walk (uid -2 - rank), placement candidates of an arrival */
};

static inline void sapio_philox4x32_10(uint32_t ctr[4], const uint32_t key_in[2]) {
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * ctr[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * ctr[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ ctr[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ ctr[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

static inline void sapio_rng4(uint64_t seed, int64_t uid, int32_t tick, uint32_t purpose, uint32_t idx, uint32_t out[4]) {
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    out[0] = (uint32_t)(uint64_t)uid; out[1] = (uint32_t)((uint64_t)uid >> 32);
    out[2] = (uint32_t)tick; out[3] = (purpose << 24) | (idx & 0xFFFFFFu);
    sapio_philox4x32_10(out, key);
}

static inline float sapio_u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

/* sequential stream: draw n -> word (n & 3) of block (n >> 2) */
typedef struct { uint64_t seed; int64_t uid; int32_t tick; uint32_t purpose; uint32_t n; } SapioStream;
static inline float sapio_stream_u(SapioStream *s) {
    uint32_t o[4];
    sapio_rng4(s->seed, s->uid, s->tick, s->purpose, s->n >> 2, o);
    float u = sapio_u01(o[s->n & 3]);
    s->n++;
    return u;
}
#endif
