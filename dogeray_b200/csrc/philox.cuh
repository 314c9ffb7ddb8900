// Counter-based sampler of the path tracer (replaces the clock()-seeded cuRAND XORWOW state of
// raygpu/kernel.cu:1061-1065 and every curand_uniform_double call, kernel.cu:644, 657, 1067-1068).
//
//   word(seed, x, y, sample, n) = Philox4x32-10(key = {seed.lo, seed.hi}, ctr = {x, y, sample, n >> 2})[n & 3]
//   uniform(n)                  = float((word >> 8) + 0.5f) * 2^-24   in (0,1] (k + 0.5 rounds in float for k >= 2^23)
//
// n is the running draw index of one (pixel, sample) path, consumed in the reference's call order
// (jitter u, jitter v, lens disk attempts, then per bounce what the material draws).  Because every
// uniform is a float on the 2^-25 grid, the reference's `u * 2 - 1` evaluated in double and
// rounded to float equals the same expression evaluated in float.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define DRB_HD __host__ __device__ __forceinline__
#else
#define DRB_HD inline
#endif

struct Philox4 { uint32_t v[4]; };

DRB_HD Philox4 philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3)
{
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

// Per-path stream.  Only `draws` has to survive between kernels of the wavefront; the four
// buffered words are regenerated on demand.
struct PathRng {
    uint32_t k0, k1, x, y, sample, draws;
    uint32_t w0, w1, w2, w3;
    DRB_HD void init(uint64_t seed, uint32_t px, uint32_t py, uint32_t s, uint32_t ndraws)
    {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32); x = px; y = py; sample = s; draws = ndraws;
        if (ndraws & 3u) refill(ndraws >> 2);
    }
    DRB_HD void refill(uint32_t block)
    {
        Philox4 p = philox4x32_10(k0, k1, x, y, sample, block);
        w0 = p.v[0]; w1 = p.v[1]; w2 = p.v[2]; w3 = p.v[3];
    }
    DRB_HD uint32_t word()
    {
        const uint32_t n = draws++;
        const uint32_t lane = n & 3u;
        if (lane == 0u) refill(n >> 2);
        return lane == 0u ? w0 : (lane == 1u ? w1 : (lane == 2u ? w2 : w3));
    }
    static DRB_HD float to_uniform(uint32_t w) { return ((float)(w >> 8) + 0.5f) * (1.0f / 16777216.0f); }
    DRB_HD float uniform() { return to_uniform(word()); }
    // The next three words at once (draws n, n+1, n+2) -- the definition of what one unit-sphere attempt consumes; k_shade's
    // pooled sampling stage (render.cu) inlines exactly this selection.  Same stream as three word() calls, but the block function
    // sits at ONE call site: lanes of a warp whose streams are at different phases (n & 3) generate their next block
    // together instead of one phase after the other, which is what three word() calls compile to.
    DRB_HD void words3(uint32_t& a, uint32_t& b, uint32_t& c)
    {
        const uint32_t ph = draws & 3u;
        uint32_t n0 = 0, n1 = 0, n2 = 0, n3 = 0;
        if (ph != 1u) {                                   // phase 1 still holds w1 w2 w3
            Philox4 p = philox4x32_10(k0, k1, x, y, sample, (draws + 3u) >> 2);
            n0 = p.v[0]; n1 = p.v[1]; n2 = p.v[2]; n3 = p.v[3];
        }
        a = ph == 0u ? n0 : (ph == 1u ? w1 : (ph == 2u ? w2 : w3));
        b = ph == 0u ? n1 : (ph == 1u ? w2 : (ph == 2u ? w3 : n0));
        c = ph == 0u ? n2 : (ph == 1u ? w3 : (ph == 2u ? n0 : n1));
        if (ph != 1u) { w0 = n0; w1 = n1; w2 = n2; w3 = n3; }
        draws += 3u;
    }
};
