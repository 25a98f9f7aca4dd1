// Blending of the 8 nearest node transforms (device functions shared by the point and voxel kernels).
#pragma once

#include "dfu_math.cuh"

namespace dfu {

// Node::getTransformationWeight for the 8 neighbours of p (missing neighbours, idx < 0, get weight 0)
DFU_DEV void neighbour_weights(const Top8& t, float px, float py, float pz, const float4* __restrict__ pos_w,
                               float (&w)[DFU_KNN]) {
#pragma unroll
    for (int k = 0; k < DFU_KNN; ++k) {
        const int j = t.i[k];
        if (j < 0) {
            w[k] = 0.f;
        } else {
            const float4 nd = __ldg(&pos_w[j]);
            w[k] = node_weight(nd.x, nd.y, nd.z, nd.w, px, py, pz, t.d[k]);
        }
    }
}

// Warpfield::calcDQB (src/dynfu/warp_field.cpp:127-148)
DFU_DEV DQ blend_ref_compose(const Top8& t, const float (&w)[DFU_KNN], const float4* __restrict__ real,
                             const float4* __restrict__ dual) {
    DQ sum = dq_from_translation(0.f, 0.f, 0.f);  // DualQuaternion(0,0,0,0,0,0), warp_field.cpp:133
#pragma unroll
    for (int k = 0; k < DFU_KNN; ++k) {
        const int j = t.i[k];
        if (j < 0) break;  // fewer than 8 nodes: the reference loops over what knnSearch returned
        const Quat nr = make_quat(__ldg(&real[j]));
        const Quat wd = qscale(make_quat(__ldg(&dual[j])), w[k]);  // operator*(T): dual only (dual_quaternion.hpp:120)
        // operator*=(DQ): dual first with the OLD real, then real (dual_quaternion.hpp:131-135)
        const Quat ndual = qadd(qmul(sum.real, wd), qmul(sum.dual, nr));
        sum.real = qmul(sum.real, nr);
        sum.dual = ndual;
    }
    // normalize(): real only (dual_quaternion.hpp:139-144)
    const float magnitude = __fsqrt_rn(qdot(sum.real, sum.real));
    sum.real = qscale(sum.real, __fdiv_rn(1.0f, magnitude));
    return sum;
}

// True dual-quaternion blending (north-star mode; no reference implementation): sign-aligned weighted
// sum, both parts divided by |real|; no support (all weights underflow to 0) -> identity.
DFU_DEV DQ blend_dqb_sum(const Top8& t, const float (&w)[DFU_KNN], const float4* __restrict__ real,
                         const float4* __restrict__ dual) {
    Quat ar{0.f, 0.f, 0.f, 0.f}, ad{0.f, 0.f, 0.f, 0.f}, r0{1.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < DFU_KNN; ++k) {
        const int j = t.i[k];
        if (j < 0) break;
        const Quat nr = make_quat(__ldg(&real[j]));
        if (k == 0) r0 = nr;
        const float ws = qdot(nr, r0) < 0.f ? -w[k] : w[k];
        ar = qadd(ar, qscale(nr, ws));
        ad = qadd(ad, qscale(make_quat(__ldg(&dual[j])), ws));
    }
    const float m2 = qdot(ar, ar);
    if (!(m2 > 0.f)) return dq_identity();
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(m2));
    return DQ{qscale(ar, inv), qscale(ad, inv)};
}

DFU_DEV DQ blend(int mode, const Top8& t, const float (&w)[DFU_KNN], const float4* __restrict__ real,
                 const float4* __restrict__ dual) {
    return mode == DFU_BLEND_REF_COMPOSE ? blend_ref_compose(t, w, real, dual) : blend_dqb_sum(t, w, real, dual);
}

// Fast path for a translation-only field (every node real == (1,0,0,0), dual.w == 0 -- the only state
// the reference ever produces, src/dynfu/utils/opt_solver.cpp:280-281).  Bit-identical to
// blend_ref_compose + dq_transform_vertex in that state (DESIGN.md derives it): the product chain
// reduces to acc += dual_k * w_k and the vertex to p + 2*acc.  Nodes whose weight is exactly 0 add +-0.
DFU_DEV V3 warp_translation_only(const Top8& t, float px, float py, float pz, const float4* __restrict__ pos_w,
                                 const float4* __restrict__ dual) {
    float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
    for (int k = 0; k < DFU_KNN; ++k) {
        const int j = t.i[k];
        if (j < 0) break;
        const float4 nd = __ldg(&pos_w[j]);
        const float w = node_weight(nd.x, nd.y, nd.z, nd.w, px, py, pz, t.d[k]);
        if (w != 0.f) {
            const float4 du = __ldg(&dual[j]);  // (w,x,y,z) stored in (x,y,z,w)
            ax = fadd(fmul(du.y, w), ax);
            ay = fadd(fmul(du.z, w), ay);
            az = fadd(fmul(du.w, w), az);
        }
    }
    return V3{fadd(px, fmul(2.f, ax)), fadd(py, fmul(2.f, ay)), fadd(pz, fmul(2.f, az))};
}

}  // namespace dfu
