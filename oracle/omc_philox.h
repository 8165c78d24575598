/*
 * omc_philox.h -- counter-based RNG shared by the CPU oracle and the instrumented reference build.
 * TEST INFRASTRUCTURE (oracle/): never linked into the product library.
 *
 * Philox4x32-10 (Salmon et al., SC'11; Random123 philox.h).  The reference's own generator is
 * RANMAR (src/omc_random.c:58-187), a sequential 97-word lagged Fibonacci state per OpenMP thread;
 * BASELINE.json's north_star replaces it by per-history counter-based streams so that results do
 * not depend on thread scheduling.  The stream layout is the GPU library's (ompmc_b200/csrc/rng.cuh):
 *
 *   key     = (seed0, seed1)                      seed0 = ixx, seed1 = jxx  ("rng seeds")
 *   counter = (block, stream, hist_lo, hist_hi)   block = draw index / 4
 *   draw k of a stream = word (k & 3) of Philox(counter with block = k >> 2), as u32 * 2^-32
 *
 * which keeps setRandom()'s contract (src/omc_random.c:172-187): a double in [0,1), 0.0 possible,
 * on a lattice at least as fine as RANMAR's 2^-24.
 */
#ifndef OMC_PHILOX_H
#define OMC_PHILOX_H
#include <stdint.h>

typedef struct omc_philox {
    uint32_t key[2];
    uint32_t ctr[4];     /* ctr[0] = next block to generate */
    uint32_t buf[4];
    uint32_t pos;        /* next unread word in buf (4 = empty) */
    uint64_t ndraws;
} omc_philox;

static inline void omc_philox_block(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline void omc_philox_seed(omc_philox *g, uint32_t seed0, uint32_t seed1, uint64_t hist, uint32_t stream) {
    g->key[0] = seed0; g->key[1] = seed1;
    g->ctr[0] = 0; g->ctr[1] = stream;
    g->ctr[2] = (uint32_t)hist; g->ctr[3] = (uint32_t)(hist >> 32);
    g->pos = 4; g->ndraws = 0;
}

static inline double omc_philox_next(omc_philox *g) {
    if (g->pos >= 4) {
        omc_philox_block(g->ctr, g->key, g->buf);
        g->ctr[0] += 1;
        g->pos = 0;
    }
    g->ndraws++;
    return (double)g->buf[g->pos++] * (1.0 / 4294967296.0);
}
#endif
