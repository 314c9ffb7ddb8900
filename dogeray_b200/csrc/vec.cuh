// float3 helpers for the tracer.  Component-wise operators and left-associated dot products: the
// same arithmetic shape as the reference's helpers (raygpu/kernel.cu:145-232), so that with
// -fmad=false every expression rounds like the reference's host-compiled code.
#pragma once
#include <cuda_runtime.h>

#define DRB_D __device__ __forceinline__

struct f3 { float x, y, z; };
DRB_D f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
DRB_D f3 mk3(float a) { return mk3(a, a, a); }
DRB_D f3 xyz(const float4& v) { return mk3(v.x, v.y, v.z); }
DRB_D f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
DRB_D f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
DRB_D f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
DRB_D f3 operator/(f3 a, f3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
DRB_D f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
DRB_D float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DRB_D f3 cross(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
DRB_D float length(f3 a) { return sqrtf(dot(a, a)); }
DRB_D f3 normalize(f3 v)
{
    const float inv = 1.0f / sqrtf(dot(v, v));          // getNormalizedVec, kernel.cu:179-183
    return mk3(v.x * inv, v.y * inv, v.z * inv);
}
