// Device arithmetic of the hot path.  Every operation is an explicitly rounded intrinsic
// (__fmul_rn/__fadd_rn/..., never contracted into FMA; FMA only where written as __fmaf_rn), in the
// evaluation order of the reference's host C++ (plain x86-64, no contraction), so results can be
// compared bit for bit with the CPU oracle.  Citations are into the reference tree.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/dynfu_b200.h"

#define DFU_DEV __device__ __forceinline__

// 256-bit global loads (sm_100: LDG.E.256) of 32-byte aligned records that are immutable while the kernel runs (8 neighbour ids or
// 8 weights of a point / voxel): one request instead of two.  `_cs`: evict-first, for records read exactly once.
DFU_DEV void ld256(const float* p, float (&r)[8]) {
    asm("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
        : "l"(p));
}
DFU_DEV void ld256(const int32_t* p, int (&r)[8]) {
    asm("ld.global.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "l"(p));
}
DFU_DEV void ld256_cs(const float* p, float (&r)[8]) {
    asm("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
        : "l"(p));
}

namespace dfu {

DFU_DEV float fmul(float a, float b) { return __fmul_rn(a, b); }
// a + b and a - b, rounded once, issued as FFMA with a unit multiplier.  fma(a, 1, b) == a + b and
// fma(b, -1, a) == a - b bit for bit (the product is exact; signed zeros, infinities and NaNs behave alike).
// Why: on sm_100 FADD/FSETP/FSEL issue on the half-rate ALU pipe while FFMA/FMUL use the full-rate FMA pipe
// (ncu: alu 67 % vs fma 14 % busy in the voxel kNN loop before this change) -- the distance / quaternion
// arithmetic is add-heavy, so the adds are moved to the FMA pipe.
DFU_DEV float fadd(float a, float b) {
    float r;
    asm("fma.rn.f32 %0, %1, 0f3F800000, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
DFU_DEV float fsub(float a, float b) {
    float r;
    asm("fma.rn.f32 %0, %1, 0fBF800000, %2;" : "=f"(r) : "f"(b), "f"(a));
    return r;
}

struct Quat {
    float w, x, y, z;
};
DFU_DEV Quat make_quat(float4 v) { return Quat{v.x, v.y, v.z, v.w}; }  // storage order (w,x,y,z) in a float4
DFU_DEV float4 to_float4(Quat q) { return make_float4(q.w, q.x, q.y, q.z); }

// boost::math::quaternion operator*= : at = a*ar-b*br-c*cr-d*dr ... evaluated left to right
DFU_DEV Quat qmul(const Quat& p, const Quat& q) {
    const float a = p.w, b = p.x, c = p.y, d = p.z;
    const float ar = q.w, br = q.x, cr = q.y, dr = q.z;
    Quat r;
    r.w = fsub(fsub(fsub(fmul(a, ar), fmul(b, br)), fmul(c, cr)), fmul(d, dr));
    r.x = fsub(fadd(fadd(fmul(a, br), fmul(b, ar)), fmul(c, dr)), fmul(d, cr));
    r.y = fadd(fadd(fsub(fmul(a, cr), fmul(b, dr)), fmul(c, ar)), fmul(d, br));
    r.z = fadd(fsub(fadd(fmul(a, dr), fmul(b, cr)), fmul(c, br)), fmul(d, ar));
    return r;
}
DFU_DEV Quat qadd(const Quat& p, const Quat& q) { return Quat{fadd(p.w, q.w), fadd(p.x, q.x), fadd(p.y, q.y), fadd(p.z, q.z)}; }
DFU_DEV Quat qscale(const Quat& p, float s) { return Quat{fmul(p.w, s), fmul(p.x, s), fmul(p.y, s), fmul(p.z, s)}; }
DFU_DEV Quat qdiv(const Quat& p, float s) { return Quat{__fdiv_rn(p.w, s), __fdiv_rn(p.x, s), __fdiv_rn(p.y, s), __fdiv_rn(p.z, s)}; }
DFU_DEV float qdot(const Quat& p, const Quat& q) {
    return fadd(fadd(fadd(fmul(p.w, q.w), fmul(p.x, q.x)), fmul(p.y, q.y)), fmul(p.z, q.z));
}

struct DQ {
    Quat real, dual;
};
DFU_DEV DQ dq_identity() { return DQ{{1.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}; }

// DualQuaternion(0,0,0, x,y,z) (include/dynfu/utils/dual_quaternion.hpp:42-67 with zero angles):
// rotation (1,0,0,0)/1, dual = ((0,t) * real) * 0.5f
DFU_DEV DQ dq_from_translation(float x, float y, float z) {
    DQ d;
    d.real = qdiv(Quat{1.f, 0.f, 0.f, 0.f}, 1.f);
    d.dual = qscale(qmul(Quat{0.f, x, y, z}, d.real), 0.5f);
    return d;
}
// dual_quaternion.hpp:127-129
DFU_DEV DQ dq_mul(const DQ& a, const DQ& b) { return DQ{qmul(a.real, b.real), qadd(qmul(a.real, b.dual), qmul(a.dual, b.real))}; }

struct V3 {
    float x, y, z;
};
DFU_DEV V3 cross(const V3& a, const V3& b) {  // cv::Vec3f::cross
    return V3{fsub(fmul(a.y, b.z), fmul(a.z, b.y)), fsub(fmul(a.z, b.x), fmul(a.x, b.z)), fsub(fmul(a.x, b.y), fmul(a.y, b.x))};
}
DFU_DEV V3 vadd(const V3& a, const V3& b) { return V3{fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)}; }
DFU_DEV V3 vsub(const V3& a, const V3& b) { return V3{fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)}; }
DFU_DEV V3 vscale(float s, const V3& a) { return V3{fmul(a.x, s), fmul(a.y, s), fmul(a.z, s)}; }

// DualQuaternion::transformVertex (dual_quaternion.hpp:204-215)
DFU_DEV V3 dq_transform_vertex(const DQ& d, const V3& v) {
    V3 rv{d.real.x, d.real.y, d.real.z};
    V3 dv{d.dual.x, d.dual.y, d.dual.z};
    V3 a = vscale(2.f, cross(rv, vadd(cross(rv, v), vscale(d.real.w, v))));
    V3 b = vscale(2.f, vadd(vsub(vscale(d.real.w, dv), vscale(d.dual.w, rv)), cross(rv, dv)));
    return vadd(vadd(v, a), b);
}
DFU_DEV V3 dq_rotate(const DQ& d, const V3& v) {
    V3 rv{d.real.x, d.real.y, d.real.z};
    V3 a = vscale(2.f, cross(rv, vadd(cross(rv, v), vscale(d.real.w, v))));
    return vadd(v, a);
}

// squared distance of nanoflann's L2_Simple_Adaptor::evalMetric (nanoflann.hpp:338-345):
// r = 0; r += dx*dx; r += dy*dy; r += dz*dz   (0 + dx*dx == dx*dx exactly)
DFU_DEV float dist2(float qx, float qy, float qz, float px, float py, float pz) {
    const float dx = fsub(qx, px), dy = fsub(qy, py), dz = fsub(qz, pz);
    return fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
}

// Node::getTransformationWeight (src/dynfu/utils/node.cpp:29-36): float differences, squares / exp in
// DOUBLE, result cast to float.  d2f = the float squared distance already known from the kNN: when
// d2f > 209*dg_w^2 the double result is < 2^-150 and rounds to exactly 0.f, so the FP64 path is skipped.
DFU_DEV float node_weight(float nx, float ny, float nz, float dg_w, float px, float py, float pz, float d2f) {
    if (d2f > fmul(209.f, fmul(dg_w, dg_w))) return 0.f;
    const double dx = (double) fsub(nx, px), dy = (double) fsub(ny, py), dz = (double) fsub(nz, pz);
    const double distSq = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const double w2 = __dmul_rn((double) dg_w, (double) dg_w);
    return __double2float_rn(exp(__ddiv_rn(-distSq, __dmul_rn(2.0, w2))));
}

// pack_tsdf / unpack_tsdf (include/kfusion/cuda/device.hpp:59-67): ushort2{half bits, weight} in a u32
DFU_DEV uint32_t pack_tsdf(float tsdf, int weight) {
    return (uint32_t) __half_as_ushort(__float2half_rn(tsdf)) | ((uint32_t) weight << 16);
}
DFU_DEV float unpack_tsdf(uint32_t v, int& weight) {
    weight = (int) (v >> 16);
    return __half2float(__ushort_as_half((unsigned short) (v & 0xffffu)));
}

// ---- sorted top-8 by (dist2, then scan order) kept in registers ---------------------------------
struct Top8 {
    float d[DFU_KNN];
    int i[DFU_KNN];
};
DFU_DEV void top8_init(Top8& t) {
#pragma unroll
    for (int k = 0; k < DFU_KNN; ++k) {
        t.d[k] = INFINITY;
        t.i[k] = -1;
    }
}
// KNNResultSet::addPoint (nanoflann.hpp:103-128): accepted only if dist < worst, placed after every
// entry with dist' <= dist.  Candidates are offered in ascending index order, so equal distances keep
// the lower index first: key (dist2, idx).
DFU_DEV void top8_insert(Top8& t, float dist, int idx) {
#pragma unroll
    for (int k = DFU_KNN - 1; k > 0; --k) {
        const bool shift = t.d[k - 1] > dist;
        const bool here = !shift && (t.d[k] > dist);
        t.i[k] = shift ? t.i[k - 1] : (here ? idx : t.i[k]);
        t.d[k] = shift ? t.d[k - 1] : (here ? dist : t.d[k]);
    }
    if (t.d[0] > dist) {
        t.d[0] = dist;
        t.i[0] = idx;
    }
}

}  // namespace dfu
