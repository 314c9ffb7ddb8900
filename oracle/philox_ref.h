/* TEST INFRASTRUCTURE -- oracle side of the sampler contract.  Not shipped, not
 * linked into the product library.
 *
 * The reference seeds cuRAND XORWOW from clock() once per sample
 * (raygpu/kernel.cu:1061-1065) and draws everything through
 * curand_uniform_double (kernel.cu:644, 657, 1067-1068), so its stream is not
 * reproducible.  Both sides of the parity tests therefore use the same
 * counter-based stream instead:
 *
 *   word(seed, x, y, sample, n) = Philox4x32-10(key = {seed.lo, seed.hi},
 *                                  ctr = {x, y, sample, n >> 2})[n & 3]
 *   uniform(n) = float((word >> 8) + 0.5f) * 2^-24   in (0, 1] (float rounding of k + 0.5 for k >= 2^23)
 *
 * n counts the draws of one (pixel, sample) path in call order, exactly the
 * order the reference consumes them.  The product has its
 * own copy of this definition in dogeray_b200/csrc/philox.cuh; tests check the
 * two against the known-answer vectors in tests/golden/philox_kat.json.
 */
#ifndef DOGERAY_ORACLE_PHILOX_REF_H
#define DOGERAY_ORACLE_PHILOX_REF_H
#include <stdint.h>

static inline void orc_philox4x32_10(uint32_t k0, uint32_t k1, const uint32_t ctr[4], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct {
    uint64_t seed;
    uint32_t x, y, sample;
    uint32_t draws;      /* number of words consumed so far */
    uint32_t buf[4];
} orc_rng;

static inline void orc_rng_init(orc_rng* r, uint64_t seed, uint32_t x, uint32_t y, uint32_t sample)
{
    r->seed = seed; r->x = x; r->y = y; r->sample = sample; r->draws = 0;
}

static inline uint32_t orc_rng_word(orc_rng* r)
{
    uint32_t n = r->draws++;
    if ((n & 3u) == 0u) {
        uint32_t ctr[4] = { r->x, r->y, r->sample, n >> 2 };
        orc_philox4x32_10((uint32_t)r->seed, (uint32_t)(r->seed >> 32), ctr, r->buf);
    }
    return r->buf[n & 3u];
}

/* (0,1] on the 2^-25 grid, like curand_uniform_double's range; 2u-1 is exact in float */
static inline float orc_rng_uniform(orc_rng* r)
{
    return ((float)(orc_rng_word(r) >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

#endif
