// oracle/ref_shim: stands in for Opt's cudaUtil.h (github.com/niessner/Opt, examples/shared/cudaUtil.h -- not vendored by
// the reference; include/kfusion/cuda/temp_utils.hpp:4 takes its float3 operators from there "to avoid importing the
// same operators twice").  Opt's header is NVIDIA's cutil_math.h; the operators below are restated from that published
// header: plain component-wise expressions, left to the compiler's default FMA contraction like the original.
// Test infrastructure only.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
// CUDA <= 8 (the reference's toolchain, Dockerfile:1) had the intrinsics `unsigned short __float2half_rn(float)` and
// `float __half2float(unsigned short)` built in (include/kfusion/cuda/device.hpp:59-67 uses them); CUDA 12 only has the
// __half-typed ones in cuda_fp16.h.  Same conversions, spelled as the PTX they compiled to.
// (cuda_fp16.h is included first so that a later include -- thrust pulls it in -- cannot collide with the macros.)
#include <cuda_fp16.h>
#ifdef __CUDACC__
// marching_cubes.cu:3-4 includes these two thrust headers AFTER this file; taking them here keeps the macros below out
// of thrust / libcu++ (their include guards make the later includes no-ops)
#include <thrust/device_ptr.h>
#include <thrust/scan.h>
#endif
#ifdef __CUDACC__
__device__ __forceinline__ unsigned short ref_shim_f2h(float f) {
    unsigned short h;
    asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(f));
    return h;
}
__device__ __forceinline__ float ref_shim_h2f(unsigned short h) {
    float f;
    asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h));
    return f;
}
__device__ __forceinline__ float ref_shim_h2f(__half h) { return ref_shim_h2f(__half_as_ushort(h)); }
#define __float2half_rn(f) ref_shim_f2h(f)
#define __half2float(h) ref_shim_h2f(h)
#endif
// Pre-Volta warp votes / shuffles without a member mask (marching_cubes.cu:81,102,113; tsdf_volume.cu:447,536;
// temp_utils.hpp:524-589): ptxas rejects them for sm_70+.  All call sites run with whole warps converged, so the
// full-mask *_sync forms are the same operation.
#define __ballot(p) __ballot_sync(0xffffffffu, (p))
#define __all(p) __all_sync(0xffffffffu, (p))
#define __any(p) __any_sync(0xffffffffu, (p))
#define __shfl_xor(v, m) __shfl_xor_sync(0xffffffffu, (v), (m))
#ifndef REF_SHIM_CUTIL_MATH
#define REF_SHIM_CUTIL_MATH
inline __host__ __device__ float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline __host__ __device__ void operator+=(float3 &a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
inline __host__ __device__ float3 operator+(float3 a, float b) { return make_float3(a.x + b, a.y + b, a.z + b); }
inline __host__ __device__ void operator+=(float3 &a, float b) { a.x += b; a.y += b; a.z += b; }
inline __host__ __device__ float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline __host__ __device__ void operator-=(float3 &a, float3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
inline __host__ __device__ float3 operator-(float3 a, float b) { return make_float3(a.x - b, a.y - b, a.z - b); }
inline __host__ __device__ float3 operator-(float3 a) { return make_float3(-a.x, -a.y, -a.z); }
inline __host__ __device__ float3 operator*(float3 a, float3 b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline __host__ __device__ float3 operator*(float3 a, float b) { return make_float3(a.x * b, a.y * b, a.z * b); }
inline __host__ __device__ float3 operator*(float b, float3 a) { return make_float3(b * a.x, b * a.y, b * a.z); }
inline __host__ __device__ float3 operator/(float3 a, float b) { return make_float3(a.x / b, a.y / b, a.z / b); }
inline __host__ __device__ float3 operator/(float b, float3 a) { return make_float3(b / a.x, b / a.y, b / a.z); }
inline __host__ __device__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline __host__ __device__ float length(float3 v) { return sqrtf(dot(v, v)); }
#endif
#ifdef __CUDACC__
#include <ref_shim_texture.h>
#endif
