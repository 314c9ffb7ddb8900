/* TEST INFRASTRUCTURE -- lets g++ compile the reference's own device + host
 * functions (a line-range slice of /root/reference/raygpu/kernel.cu, piped in by
 * oracle/make_ref.py; the slice itself is never written into this repo) as plain
 * host C++.  Nothing here restates reference logic: it only supplies the CUDA
 * vocabulary the slice expects.
 *
 *   __device__/__host__/__global__        -> nothing
 *   __constant__                          -> static
 *   float3/float4/uchar4/make_*           -> plain structs
 *   int3                                  -> members that remember the float they
 *                                            were assigned, so Kernel()'s
 *                                            `outputr[w].x = Color*255*scale`
 *                                            (kernel.cu:1083-1085) can be read back
 *                                            unquantised
 *   blockIdx/blockDim/threadIdx/gridDim   -> thread_local
 *   tex2D<uchar4>                         -> point sample, wrap, normalised coords
 *                                            (the texture descriptor of
 *                                            kernel.cu:1959-1964)
 *   curandState/curand_init/
 *   curand_uniform_double                 -> the Philox stream of philox_ref.h
 *   SDL message box, GetTickCount         -> stderr / constant
 */
#ifndef DOGERAY_ORACLE_REF_HOST_SHIM_H
#define DOGERAY_ORACLE_REF_HOST_SHIM_H

#include <cmath>
#include <math.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <ctime>
#include <iostream>
#include <fstream>
#include <sstream>
#include <string>
#include <algorithm>
#include <vector>
#include "philox_ref.h"

#define __device__
#define __host__
#define __global__
#define __constant__ static

struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
static inline float3 make_float3(float x, float y, float z) { float3 r; r.x = x; r.y = y; r.z = z; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

/* an int that remembers the float it was assigned */
struct orc_capint {
    int i; float f;
    orc_capint() : i(0), f(0.0f) {}
    orc_capint& operator=(int v) { i = v; f = (float)v; return *this; }
    orc_capint& operator=(float v)
    {
        f = v;
        /* device float->int is cvt.rzi: NaN -> 0, saturating */
        if (v != v) i = 0;
        else if (v >= 2147483648.0f) i = INT32_MAX;
        else if (v <= -2147483648.0f) i = INT32_MIN;
        else i = (int)v;
        return *this;
    }
    orc_capint& operator=(double v) { return (*this = (float)v); }
    operator int() const { return i; }
};
struct int3 { orc_capint x, y, z; };

struct orc_dim3 { unsigned x, y, z; };
static thread_local orc_dim3 blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, threadIdx = {0, 0, 0}, gridDim = {1, 1, 1};

typedef int cudaError_t;
typedef unsigned long long cudaTextureObject_t;     /* index into orc_textures */

struct orc_texture { int w, h; std::vector<unsigned char> rgba; };
static std::vector<orc_texture> orc_textures;

template <typename T> static inline T tex2D(cudaTextureObject_t t, float u, float v);
template <> inline uchar4 tex2D<uchar4>(cudaTextureObject_t t, float u, float v)
{
    uchar4 r = {0, 0, 0, 0};
    if (t >= orc_textures.size()) return r;
    const orc_texture& tx = orc_textures[(size_t)t];
    if (tx.w <= 0 || tx.h <= 0) return r;
    float fu = u - floorf(u), fv = v - floorf(v);
    int ix = (int)floorf(fu * (float)tx.w), iy = (int)floorf(fv * (float)tx.h);
    if (!(ix >= 0)) ix = 0;
    if (!(iy >= 0)) iy = 0;
    if (ix >= tx.w) ix = tx.w - 1;
    if (iy >= tx.h) iy = tx.h - 1;
    const unsigned char* p = &tx.rgba[4 * ((size_t)iy * tx.w + ix)];
    r.x = p[0]; r.y = p[1]; r.z = p[2]; r.w = p[3];
    return r;
}

/* sampler: which (pixel, sample) a state belongs to comes from the thread-locals
 * the driver sets before calling Kernel()/raycolor() */
static uint64_t orc_seed = 0;
static thread_local uint32_t orc_next_sample = 0;
static thread_local uint64_t orc_rays = 0;           /* not used by the slice; driver statistic */
typedef orc_rng curandState;
static inline void curand_init(unsigned long long, unsigned long long, unsigned long long, curandState* s)
{
    uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t y = blockIdx.y * blockDim.y + threadIdx.y;
    orc_rng_init(s, orc_seed, x, y, orc_next_sample++);
}
static inline double curand_uniform_double(curandState* s) { return (double)orc_rng_uniform(s); }

/* CUDA's mixed-type min used at kernel.cu:679 and :920 */
static inline double min(float a, double b) { return fmin((double)a, b); }
static inline double min(double a, float b) { return fmin(a, (double)b); }

#define SDL_MESSAGEBOX_ERROR 0
#define SDL_ShowSimpleMessageBox(flags, title, msg, win) fprintf(stderr, "[ref] %s: %s\n", title, msg)
static inline unsigned GetTickCount() { return 12345u; }

#endif
