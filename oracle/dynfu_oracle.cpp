/*
 * dynfu_oracle.cpp -- CPU ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * A dependency-free C++17 restatement of the arithmetic of swarth100/dynfu's per-frame hot path.
 * Compile with -ffp-contract=off: the reference is plain x86-64 C++ (no FMA contraction), so every
 * float operation below is rounded separately, in the reference's evaluation order.
 *
 * Citations are into /root/reference.  Quaternions are (w,x,y,z), Hamilton product.
 *
 * Built twice by oracle/Makefile:
 *   oracle/libdynfu_oracle.so          kNN = brute force, key (dist2, idx)          [default checker]
 *   oracle/_ref/libdynfu_oracle_nf.so  kNN = the reference's vendored nanoflann KD-tree
 *                                      (-DORC_USE_NANOFLANN, header compiled where it lies)
 */
#include "dynfu_oracle.h"

#include <algorithm>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORC_USE_NANOFLANN
#include <nanoflann.hpp> /* /root/reference/include/nanoflann/nanoflann.hpp, v0x123 */
#endif

namespace {

constexpr int KNN = 8; /* include/dynfu/warp_field.hpp:27 */

/* ------------------------------------------------------------------------------------------ */
/* boost::math::quaternion<float> subset                                                       */

struct Quat {
    float w, x, y, z;
};

/* boost/math/quaternion.hpp operator*= : at = a*ar-b*br-c*cr-d*dr, ... evaluated left to right */
inline Quat qmul(const Quat& p, const Quat& q) {
    const float a = p.w, b = p.x, c = p.y, d = p.z;
    const float ar = q.w, br = q.x, cr = q.y, dr = q.z;
    Quat r;
    r.w = a * ar - b * br - c * cr - d * dr;
    r.x = a * br + b * ar + c * dr - d * cr;
    r.y = a * cr - b * dr + c * ar + d * br;
    r.z = a * dr + b * cr - c * br + d * ar;
    return r;
}
inline Quat qadd(const Quat& p, const Quat& q) { return {p.w + q.w, p.x + q.x, p.y + q.y, p.z + q.z}; }
inline Quat qsub(const Quat& p, const Quat& q) { return {p.w - q.w, p.x - q.x, p.y - q.y, p.z - q.z}; }
inline Quat qscale(const Quat& p, float s) { return {p.w * s, p.x * s, p.y * s, p.z * s}; }
inline Quat qdiv(const Quat& p, float s) { return {p.w / s, p.x / s, p.y / s, p.z / s}; }
inline Quat qconj(const Quat& p) { return {p.w, -p.x, -p.y, -p.z}; }
/* boost::math::norm(q) is the Cayley norm = SUM OF SQUARES (real(q*conj(q))) */
inline float qnorm_boost(const Quat& p) { return p.w * p.w + p.x * p.x + p.y * p.y + p.z * p.z; }

struct DQ {
    Quat real, dual;
};

inline DQ load_dq(const float* a) { return {{a[0], a[1], a[2], a[3]}, {a[4], a[5], a[6], a[7]}}; }
inline void store_dq(const DQ& d, float* o) {
    o[0] = d.real.w; o[1] = d.real.x; o[2] = d.real.y; o[3] = d.real.z;
    o[4] = d.dual.w; o[5] = d.dual.x; o[6] = d.dual.y; o[7] = d.dual.z;
}

/* dual_quaternion.hpp:31,42-45 : real = rot / norm(rot) [squared norm!], dual = ((0,t)*real)*0.5f */
inline DQ dq_from_rot_trans(const Quat& rot, const float t[3]) {
    DQ d;
    d.real = qdiv(rot, qnorm_boost(rot));
    d.dual = qscale(qmul(Quat{0.f, t[0], t[1], t[2]}, d.real), 0.5f);
    return d;
}

/* dual_quaternion.hpp:48-67 : cos/sin of (T * 0.5) are evaluated in double, stored to T */
inline DQ dq_from_euler(float yaw, float pitch, float roll, float x, float y, float z) {
    float cy = (float) std::cos(yaw * 0.5);
    float sy = (float) std::sin(yaw * 0.5);
    float cr = (float) std::cos(roll * 0.5);
    float sr = (float) std::sin(roll * 0.5);
    float cp = (float) std::cos(pitch * 0.5);
    float sp = (float) std::sin(pitch * 0.5);
    float qw = cy * cr * cp + sy * sr * sp;
    float qx = cy * sr * cp - sy * cr * sp;
    float qy = cy * cr * sp + sy * sr * cp;
    float qz = sy * cr * cp - cy * sr * sp;
    float t[3] = {x, y, z};
    return dq_from_rot_trans(Quat{qw, qx, qy, qz}, t);
}

/* dual_quaternion.hpp:127-129 */
inline DQ dq_mul(const DQ& a, const DQ& b) {
    return {qmul(a.real, b.real), qadd(qmul(a.real, b.dual), qmul(a.dual, b.real))};
}

/* dual_quaternion.hpp:139-144 : only the real part is normalised */
inline bool dq_normalize(DQ& d) {
    float magnitude = sqrtf(d.real.w * d.real.w + d.real.x * d.real.x + d.real.y * d.real.y + d.real.z * d.real.z);
    bool ok = magnitude > 1.192092896e-07f;
    d.real = qscale(d.real, 1.0f / magnitude);
    return ok;
}

struct V3 {
    float x, y, z;
};
inline V3 cross(const V3& a, const V3& b) { /* cv::Vec3f::cross */
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline V3 vadd(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 vsub(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 vscale(float s, const V3& a) { return {a.x * s, a.y * s, a.z * s}; }

/* dual_quaternion.hpp:204-215
 *   vect + 2.f * rv.cross(rv.cross(vect) + rw * vect) + 2.f * (rw * dv - dw * rv + rv.cross(dv)) */
inline V3 dq_transform_vertex(const DQ& d, const V3& v) {
    V3 rv{d.real.x, d.real.y, d.real.z};
    V3 dv{d.dual.x, d.dual.y, d.dual.z};
    V3 a = vscale(2.f, cross(rv, vadd(cross(rv, v), vscale(d.real.w, v))));
    V3 b = vscale(2.f, vadd(vsub(vscale(d.real.w, dv), vscale(d.dual.w, rv)), cross(rv, dv)));
    return vadd(vadd(v, a), b);
}
/* rotation only: what a normal transform should be (not what the reference does) */
inline V3 dq_rotate(const DQ& d, const V3& v) {
    V3 rv{d.real.x, d.real.y, d.real.z};
    V3 a = vscale(2.f, cross(rv, vadd(cross(rv, v), vscale(d.real.w, v))));
    return vadd(v, a);
}

/* src/dynfu/utils/node.cpp:29-36 : float differences, pow/exp in double, result cast to float */
inline float node_weight(const float* np, float dg_w, const float* p) {
    double dx = (double) (np[0] - p[0]);
    double dy = (double) (np[1] - p[1]);
    double dz = (double) (np[2] - p[2]);
    double distSq = dx * dx + dy * dy + dz * dz;
    double w2 = (double) dg_w * (double) dg_w;
    return (float) std::exp(-distSq / (2 * w2));
}

/* ------------------------------------------------------------------------------------------ */
/* kNN index                                                                                   */

/* nanoflann L2_Simple_Adaptor::evalMetric (include/nanoflann/nanoflann.hpp:338-345) */
inline float dist2(const float* q, const float* p) {
    float r = 0.f;
    for (int i = 0; i < 3; ++i) {
        const float diff = q[i] - p[i];
        r += diff * diff;
    }
    return r;
}

#ifdef ORC_USE_NANOFLANN
/* adaptor with the interface of include/nanoflann/pointcloud.hpp:11-44 over a flat xyz array */
struct FlatCloud {
    const float* pts;
    size_t n;
    inline size_t kdtree_get_point_count() const { return n; }
    inline float kdtree_get_pt(const size_t idx, int dim) const { return pts[idx * 3 + dim]; }
    template <class BBOX>
    bool kdtree_get_bbox(BBOX&) const { return false; }
};
typedef nanoflann::L2_Simple_Adaptor<float, FlatCloud> nf_adaptor;               /* warp_field.hpp:29 */
typedef nanoflann::KDTreeSingleIndexAdaptor<nf_adaptor, FlatCloud, 3> nf_tree_t; /* warp_field.hpp:30 */
#endif

struct KnnIndex {
    const float* nodes;
    int N;
#ifdef ORC_USE_NANOFLANN
    FlatCloud cloud;
    std::unique_ptr<nf_tree_t> tree;
#endif
    KnnIndex(const float* nodes_, int N_) : nodes(nodes_), N(N_) {
#ifdef ORC_USE_NANOFLANN
        cloud.pts = nodes;
        cloud.n = (size_t) N;
        /* warp_field.cpp:26-27 : leaf size 10 */
        tree.reset(new nf_tree_t(3, cloud, nanoflann::KDTreeSingleIndexAdaptorParams(10)));
        tree->buildIndex();
#endif
    }
    /* returns n found (min(k,N)); ascending */
    int query(const float* q, int k, int32_t* idx, float* d2) const {
#ifdef ORC_USE_NANOFLANN
        /* warp_field.cpp:111-122 */
        size_t ret[16];
        float dd[16];
        int n = (int) tree->knnSearch(q, (size_t) k, ret, dd);
        for (int i = 0; i < n; ++i) {
            idx[i] = (int32_t) ret[i];
            if (d2) d2[i] = dd[i];
        }
        return n;
#else
        float bd[17];
        int32_t bi[17];
        int cnt = 0;
        for (int j = 0; j < N; ++j) {
            float d = dist2(q, nodes + 3 * (size_t) j);
            if (cnt == k && !(d < bd[k - 1])) continue; /* j ascending: equal dist keeps lower idx */
            int i = cnt < k ? cnt : k - 1;
            while (i > 0 && bd[i - 1] > d) {
                bd[i] = bd[i - 1];
                bi[i] = bi[i - 1];
                --i;
            }
            bd[i] = d;
            bi[i] = j;
            if (cnt < k) ++cnt;
        }
        for (int i = 0; i < cnt; ++i) {
            idx[i] = bi[i];
            if (d2) d2[i] = bd[i];
        }
        return cnt;
#endif
    }
};

/* src/dynfu/warp_field.cpp:127-148 (REF_COMPOSE) and the north-star's true DQB (DQB_SUM) */
inline DQ blend_point(const KnnIndex& index, const float* pos, const float* dq, const float* dg_w, const float* p,
                      int blend_mode) {
    int32_t nb[KNN];
    int n = index.query(p, KNN, nb, nullptr);
    if (blend_mode == ORC_BLEND_REF_COMPOSE) {
        DQ sum = dq_from_euler(0.f, 0.f, 0.f, 0.f, 0.f, 0.f); /* warp_field.cpp:133 */
        for (int k = 0; k < n; ++k) {
            const int j = nb[k];
            float w = node_weight(pos + 3 * (size_t) j, dg_w[j], p);
            DQ node = load_dq(dq + 8 * (size_t) j);
            DQ weighted{node.real, qscale(node.dual, w)}; /* dual_quaternion.hpp:120 : dual only */
            /* dual_quaternion.hpp:131-135 : dual first with the OLD real, then real */
            Quat nd = qadd(qmul(sum.real, weighted.dual), qmul(sum.dual, weighted.real));
            sum.real = qmul(sum.real, weighted.real);
            sum.dual = nd;
        }
        dq_normalize(sum);
        return sum;
    }
    /* DQB_SUM: Q = sum_k w_k * sign_k * q_k ; both parts divided by |real| ; no support -> identity */
    Quat ar{0.f, 0.f, 0.f, 0.f}, ad{0.f, 0.f, 0.f, 0.f};
    Quat r0{1.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < n; ++k) {
        const int j = nb[k];
        float w = node_weight(pos + 3 * (size_t) j, dg_w[j], p);
        DQ node = load_dq(dq + 8 * (size_t) j);
        if (k == 0) r0 = node.real;
        float dot = node.real.w * r0.w + node.real.x * r0.x + node.real.y * r0.y + node.real.z * r0.z;
        float ws = dot < 0.f ? -w : w;
        ar = qadd(ar, qscale(node.real, ws));
        ad = qadd(ad, qscale(node.dual, ws));
    }
    float m2 = ar.w * ar.w + ar.x * ar.x + ar.y * ar.y + ar.z * ar.z;
    if (!(m2 > 0.f)) return DQ{{1.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    float inv = 1.0f / sqrtf(m2);
    return DQ{qscale(ar, inv), qscale(ad, inv)};
}

inline float half2float(uint16_t h) {
    _Float16 f;
    std::memcpy(&f, &h, 2);
    return (float) f;
}
inline uint16_t float2half(float v) {
    _Float16 f = (_Float16) v; /* round to nearest even, like __float2half_rn */
    uint16_t h;
    std::memcpy(&h, &f, 2);
    return h;
}

inline float tukey(float tukey_offset, float c, const float e[3]) {
    /* opt_solver.cpp:204-212 : sqrt/pow promote to double */
    double s = std::sqrt((double) (e[0] * e[0] + e[1] * e[1] + e[2] * e[2])) / tukey_offset;
    if (s < c) return (float) std::pow(1.f - std::pow(s, 2) / std::pow((double) c, 2), 2);
    return 0.f;
}

} /* namespace */

/* ============================================================================================ */
extern "C" {

void orc_dq_from_rot_trans(const float rot[4], const float t[3], float out[8]) {
    store_dq(dq_from_rot_trans(Quat{rot[0], rot[1], rot[2], rot[3]}, t), out);
}
void orc_dq_from_euler(float yaw, float pitch, float roll, float x, float y, float z, float out[8]) {
    store_dq(dq_from_euler(yaw, pitch, roll, x, y, z), out);
}
/* dual_quaternion.hpp:70-86 */
void orc_dq_from_rodrigues(const float rod[3], const float t[3], float out[8]) {
    double nrm = std::sqrt((double) rod[0] * rod[0] + (double) rod[1] * rod[1] + (double) rod[2] * rod[2]);
    double theta = 2 * std::atan(nrm);
    double ith = 1.0 / theta;
    float axis[3] = {(float) (rod[0] * ith), (float) (rod[1] * ith), (float) (rod[2] * ith)};
    double an = std::sqrt((double) axis[0] * axis[0] + (double) axis[1] * axis[1] + (double) axis[2] * axis[2]);
    double ian = 1.0 / an;
    float axn[3] = {(float) (axis[0] * ian), (float) (axis[1] * ian), (float) (axis[2] * ian)};
    double s = std::sin(0.5 * theta);
    Quat rot{(float) std::cos(0.5 * theta), (float) (s * axn[0]), (float) (s * axn[1]), (float) (s * axn[2])};
    rot = qdiv(rot, qnorm_boost(rot));
    store_dq(dq_from_rot_trans(rot, t), out);
}
void orc_dq_add(const float a[8], const float b[8], float out[8]) { /* :99-101 */
    DQ x = load_dq(a), y = load_dq(b);
    store_dq(DQ{qadd(x.real, y.real), qadd(x.dual, y.dual)}, out);
}
void orc_dq_sub(const float a[8], const float b[8], float out[8]) { /* :109-111 */
    DQ x = load_dq(a), y = load_dq(b);
    store_dq(DQ{qsub(x.real, y.real), qsub(x.dual, y.dual)}, out);
}
void orc_dq_scale(const float a[8], float s, float out[8]) { /* :120-125 : dual only */
    DQ x = load_dq(a);
    store_dq(DQ{x.real, qscale(x.dual, s)}, out);
}
void orc_dq_mul(const float a[8], const float b[8], float out[8]) { store_dq(dq_mul(load_dq(a), load_dq(b)), out); }
void orc_dq_conj(const float a[8], float out[8]) { /* :137 */
    DQ x = load_dq(a);
    store_dq(DQ{qconj(x.real), qconj(x.dual)}, out);
}
int orc_dq_normalize(const float a[8], float out[8]) {
    DQ x = load_dq(a);
    bool ok = dq_normalize(x);
    store_dq(x, out);
    return ok ? 0 : 1;
}
void orc_dq_get_translation(const float a[8], float t[3]) { /* :94-97 */
    DQ x = load_dq(a);
    Quat q = qmul(qscale(x.dual, 2.0f), qconj(x.real));
    t[0] = q.x; t[1] = q.y; t[2] = q.z;
}
float orc_dq_get_roll(const float a[8]) { /* :148-161 */
    DQ d = load_dq(a);
    float sinr = (float) (+2.0 * (d.real.w * d.real.x + d.real.y * d.real.z));
    float cosr = (float) (+1.0 - 2.0 * (d.real.x * d.real.x + d.real.y * d.real.y));
    float roll = atan2f(sinr, cosr);
    if (roll > M_PI) roll -= (float) M_PI_2;
    return roll;
}
float orc_dq_get_pitch(const float a[8]) { /* :163-177 */
    DQ d = load_dq(a);
    float sinp = (float) (+2.0 * (d.real.w * d.real.y - d.real.z * d.real.x));
    if (std::fabs(sinp) >= 1) return (float) std::copysign(M_PI / 2, (double) sinp);
    return asinf(sinp);
}
float orc_dq_get_yaw(const float a[8]) { /* :179-192 */
    DQ d = load_dq(a);
    float siny = (float) (+2.0 * (d.real.w * d.real.z + d.real.x * d.real.y));
    float cosy = (float) (+1.0 - 2.0 * (d.real.y * d.real.y + d.real.z * d.real.z));
    float yaw = atan2f(siny, cosy);
    if (yaw > M_PI) yaw -= (float) M_PI_2;
    return yaw;
}
void orc_dq_get_rodrigues(const float a[8], float rod[3]) { /* :196-202 */
    DQ d = load_dq(a);
    double nrm = std::sqrt((double) d.real.x * d.real.x + (double) d.real.y * d.real.y + (double) d.real.z * d.real.z);
    float theta = 2 * acosf(d.real.w);
    double tn = std::tan(0.5 * theta);
    float q[3] = {(float) (d.real.x * tn), (float) (d.real.y * tn), (float) (d.real.z * tn)};
    double inv = 1.0 / nrm;
    rod[0] = (float) (q[0] * inv); rod[1] = (float) (q[1] * inv); rod[2] = (float) (q[2] * inv);
}
void orc_dq_transform_vertex(const float a[8], const float v[3], float out[3]) {
    V3 r = dq_transform_vertex(load_dq(a), V3{v[0], v[1], v[2]});
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
/* dual_quaternion.hpp:217-228 : the reference applies the vertex formula (translation included) */
void orc_dq_transform_normal(const float a[8], const float n[3], int normal_mode, float out[3]) {
    V3 r = normal_mode == ORC_NORMAL_REF ? dq_transform_vertex(load_dq(a), V3{n[0], n[1], n[2]})
                                         : dq_rotate(load_dq(a), V3{n[0], n[1], n[2]});
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
int orc_dq_to_string(const float a[8], char* buf, int buflen) { /* :230-232 + boost operator<< */
    return snprintf(buf, (size_t) buflen, "real: (%g,%g,%g,%g)\ndual: (%g,%g,%g,%g)\n", a[0], a[1], a[2], a[3], a[4],
                    a[5], a[6], a[7]);
}

float orc_node_weight(const float node_pos[3], float dg_w, const float p[3]) { return node_weight(node_pos, dg_w, p); }

long orc_knn(const float* nodes_xyz, int N, const float* q_xyz, long Q, int k, int32_t* idx_out,
             float* dist2_out_or_null) {
    if (k > 16) k = 16;
    KnnIndex index(nodes_xyz, N);
    long ties = 0;
#pragma omp parallel for schedule(static) reduction(+ : ties)
    for (long i = 0; i < Q; ++i) {
        int32_t idx[17];
        float d2[17];
        int kk = std::min(k + 1, N);
#ifdef ORC_USE_NANOFLANN
        kk = std::min(k, N);
#endif
        int n = index.query(q_xyz + 3 * i, kk, idx, d2);
        for (int j = 0; j + 1 < n; ++j)
            if (d2[j] == d2[j + 1]) {
                ++ties;
                break;
            }
        for (int j = 0; j < k; ++j) {
            idx_out[i * k + j] = j < n ? idx[j] : -1;
            if (dist2_out_or_null) dist2_out_or_null[i * k + j] = j < n ? d2[j] : INFINITY;
        }
    }
    return ties;
}

void orc_blend(const float* pos, const float* dq, const float* dg_w, int N, const float* p_xyz, long Q,
               int blend_mode, float* dq_out) {
    KnnIndex index(pos, N);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < Q; ++i) store_dq(blend_point(index, pos, dq, dg_w, p_xyz + 3 * i, blend_mode), dq_out + 8 * i);
}

/* src/dynfu/warp_field.cpp:150-171 */
void orc_warp(const float* pos, const float* dq, const float* dg_w, int N, const float* v, const float* n_or_null,
              long P, int blend_mode, int normal_mode, float* v_out, float* n_out_or_null) {
    KnnIndex index(pos, N);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < P; ++i) {
        DQ b = blend_point(index, pos, dq, dg_w, v + 3 * i, blend_mode);
        V3 r = dq_transform_vertex(b, V3{v[3 * i], v[3 * i + 1], v[3 * i + 2]});
        v_out[3 * i] = r.x; v_out[3 * i + 1] = r.y; v_out[3 * i + 2] = r.z;
        if (n_or_null && n_out_or_null) {
            V3 nn{n_or_null[3 * i], n_or_null[3 * i + 1], n_or_null[3 * i + 2]};
            V3 rn = normal_mode == ORC_NORMAL_REF ? dq_transform_vertex(b, nn) : dq_rotate(b, nn);
            n_out_or_null[3 * i] = rn.x; n_out_or_null[3 * i + 1] = rn.y; n_out_or_null[3 * i + 2] = rn.z;
        }
    }
}

/* src/kfusion/cuda/imgproc.cu:233-245 (called with finv = 1/f, :252) */
void orc_compute_dists(const uint16_t* depth, size_t depth_pitch_bytes, uint16_t* dists, size_t dists_pitch_bytes,
                       int rows, int cols, const float intr[4]) {
    const float finvx = 1.f / intr[0], finvy = 1.f / intr[1], cx = intr[2], cy = intr[3];
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; ++y) {
        const uint16_t* drow = (const uint16_t*) ((const char*) depth + (size_t) y * depth_pitch_bytes);
        uint16_t* orow = (uint16_t*) ((char*) dists + (size_t) y * dists_pitch_bytes);
        for (int x = 0; x < cols; ++x) {
            float xl = ((float) x - cx) * finvx;
            float yl = ((float) y - cy) * finvy;
            float lambda = sqrtf(xl * xl + yl * yl + 1);
            orow[x] = float2half((float) drow[x] * lambda * 0.001f);
        }
    }
}

/* src/kfusion/cuda/imgproc.cu:187-215 (points_normals_kernel) with Reprojector::operator()
 * (include/kfusion/cuda/device.hpp:50-54).  points/normals: rows*cols float4, NaN where invalid.
 * normalized() (include/kfusion/cuda/temp_utils.hpp:91) is v * rsqrt(dot) with the GPU's approximate rsqrt; the
 * restatement divides by the IEEE square root (documented deviation, <= 2 ulp). */
void orc_points_normals(const uint16_t* depth, size_t depth_pitch_bytes, int rows, int cols, const float intr[4],
                        float* points4, float* normals4) {
    const float finvx = 1.f / intr[0], finvy = 1.f / intr[1], cx = intr[2], cy = intr[3];
    const float qnan = std::numeric_limits<float>::quiet_NaN();
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; ++y) {
        const uint16_t* r0 = (const uint16_t*) ((const char*) depth + (size_t) y * depth_pitch_bytes);
        const uint16_t* r1 = (const uint16_t*) ((const char*) depth + (size_t) (y + 1 < rows ? y + 1 : y) * depth_pitch_bytes);
        for (int x = 0; x < cols; ++x) {
            float* P = points4 + 4 * ((size_t) y * cols + x);
            float* Nn = normals4 + 4 * ((size_t) y * cols + x);
            for (int c = 0; c < 4; ++c) P[c] = Nn[c] = qnan;
            if (x >= cols - 1 || y >= rows - 1) continue;
            const float z00 = (float) r0[x] * 0.001f, z01 = (float) r0[x + 1] * 0.001f, z10 = (float) r1[x] * 0.001f;
            if (z00 * z01 * z10 != 0) {
                auto reproj = [&](int u, int v, float z, float o[3]) {
                    o[0] = z * ((float) u - cx) * finvx;
                    o[1] = z * ((float) v - cy) * finvy;
                    o[2] = z;
                };
                float v00[3], v01[3], v10[3];
                reproj(x, y, z00, v00);
                reproj(x + 1, y, z01, v01);
                reproj(x, y + 1, z10, v10);
                const float a[3] = {v01[0] - v00[0], v01[1] - v00[1], v01[2] - v00[2]};
                const float b[3] = {v10[0] - v00[0], v10[1] - v00[1], v10[2] - v00[2]};
                const float c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
                const float len = sqrtf((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
                Nn[0] = -(c[0] / len); Nn[1] = -(c[1] / len); Nn[2] = -(c[2] / len); Nn[3] = 0.f;
                P[0] = v00[0]; P[1] = v00[1]; P[2] = v00[2]; P[3] = 0.f;
            }
        }
    }
}

/* The valid entries of a points(/normals) image in raster order, as the reference collects them on the host
 * (src/dynfu/dyn_fusion.cpp:120-134 pushes every downloaded vertex; src/dynfu/utils/frame.cpp keeps them as
 * given); xform = {R row-major, t} or NULL.  Returns the number of valid pixels. */
long orc_compact_points(const float* points4, const float* normals4_or_null, int rows, int cols, const float* xform,
                        float* out_v, float* out_n_or_null) {
    long n = 0;
    for (long i = 0; i < (long) rows * cols; ++i) {
        const float* p = points4 + 4 * i;
        const float* q = normals4_or_null ? normals4_or_null + 4 * i : nullptr;
        if (p[0] != p[0] || p[1] != p[1] || p[2] != p[2]) continue;
        if (q && (q[0] != q[0] || q[1] != q[1] || q[2] != q[2])) continue;
        float v[3] = {p[0], p[1], p[2]}, m[3] = {q ? q[0] : 0.f, q ? q[1] : 0.f, q ? q[2] : 0.f};
        if (xform) {
            for (int r = 0; r < 3; ++r) {
                v[r] = fmaf(xform[3 * r + 2], p[2], fmaf(xform[3 * r + 1], p[1], fmaf(xform[3 * r], p[0], xform[9 + r])));
                if (q) m[r] = fmaf(xform[3 * r + 2], q[2], fmaf(xform[3 * r + 1], q[1], xform[3 * r] * q[0]));
            }
        }
        for (int c = 0; c < 3; ++c) {
            out_v[3 * n + c] = v[c];
            if (out_n_or_null) out_n_or_null[3 * n + c] = m[c];
        }
        ++n;
    }
    return n;
}

/* DynFusion::findCorrespondingFrame (src/dynfu/dyn_fusion.cpp:212-242): a KD-tree over the canonical vertices,
 * 1-NN per live vertex, gather vertex and normal.  Returns the number of live vertices whose two nearest
 * canonical vertices are at bit-equal distance (0 when built against nanoflann, which cannot tell). */
long orc_find_corresponding(const float* canon_v, const float* canon_n_or_null, int P_canon, const float* live_v, long P_live,
                            float* out_v, float* out_n_or_null, int32_t* idx_out_or_null) {
    KnnIndex index(canon_v, P_canon);
    long ties = 0;
#pragma omp parallel for schedule(static) reduction(+ : ties)
    for (long i = 0; i < P_live; ++i) {
        int32_t idx[2];
        float d2[2];
#ifdef ORC_USE_NANOFLANN
        index.query(live_v + 3 * i, 1, idx, d2);
#else
        int n = index.query(live_v + 3 * i, std::min(2, P_canon), idx, d2);
        if (n == 2 && d2[0] == d2[1]) ++ties;
#endif
        for (int c = 0; c < 3; ++c) {
            out_v[3 * i + c] = canon_v[3 * (size_t) idx[0] + c];
            if (canon_n_or_null && out_n_or_null) out_n_or_null[3 * i + c] = canon_n_or_null[3 * (size_t) idx[0] + c];
        }
        if (idx_out_or_null) idx_out_or_null[i] = idx[0];
    }
    return ties;
}

/* ---- Warpfield::update (src/dynfu/warp_field.cpp:34-95) ------------------------------------------------------ */

/* getUnsupportedVertices (:34-62): a vertex is unsupported when min_k dist(v, n_k) / dg_w_k >= 1 over its 8 nearest
 * nodes.  `sqrt(pow(dx,2)+pow(dy,2)+pow(dz,2))` is evaluated in double (std::pow(float,int) promotes) and stored in a
 * float.  flags_out[P]; returns the number of unsupported vertices. */
long orc_unsupported(const float* pos, const float* dg_w, int N, const float* verts, long P, uint8_t* flags_out) {
    KnnIndex index(pos, N);
    long count = 0;
#pragma omp parallel for schedule(static) reduction(+ : count)
    for (long i = 0; i < P; ++i) {
        int32_t idx[16];
        float d2[16];
        const float* v = verts + 3 * i;
        int n = index.query(v, std::min(8, N), idx, d2);
        float mn = HUGE_VALF;
        for (int k = 0; k < n; ++k) {
            const float* c = pos + 3 * (size_t) idx[k];
            const float dx = v[0] - c[0], dy = v[1] - c[1], dz = v[2] - c[2];
            const double s = ((double) dx * (double) dx + (double) dy * (double) dy) + (double) dz * (double) dz;
            const float dist = (float) std::sqrt(s);
            const float r = dist / dg_w[idx[k]];
            if (r <= mn) mn = r;
        }
        flags_out[i] = mn >= 1 ? 1 : 0;
        count += flags_out[i];
    }
    return count;
}

/* pcl::VoxelGrid<pcl::PointXYZ>::applyFilter -- PCL 1.8.1 (CMakeLists.txt:63 of the reference; PCL is not vendored in
 * the reference tree, so this restates filters/include/pcl/filters/impl/voxel_grid.hpp:212-437 from its published
 * algorithm): bounding box -> integer cell of every point -> points grouped by cell -> one centroid per non-empty
 * cell, cells in ascending linear index.  The centroid accumulates in float (AccumulatorXYZ) and divides by the
 * count.  PCL groups with an UNSTABLE std::sort, so the order of the float additions inside a cell is
 * implementation defined there; order_mode 0 adds in ascending point index (the canonical order the GPU path
 * reproduces), order_mode 1 reproduces std::sort on (cell) keys with this libstdc++ to measure the difference.
 * Returns the number of output points, or -1 if the grid would overflow int32 (PCL then returns its input). */
long orc_voxel_grid(const float* pts, long U, const float leaf[3], float* out, int order_mode) {
    if (U <= 0) return 0;
    const float inv[3] = {1.f / leaf[0], 1.f / leaf[1], 1.f / leaf[2]};
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (long i = 0; i < U; ++i)
        for (int c = 0; c < 3; ++c) {
            mn[c] = std::min(mn[c], pts[3 * i + c]);
            mx[c] = std::max(mx[c], pts[3 * i + c]);
        }
    int64_t d[3];
    for (int c = 0; c < 3; ++c) d[c] = (int64_t) ((mx[c] - mn[c]) * inv[c]) + 1;
    if (d[0] * d[1] * d[2] > (int64_t) INT32_MAX) return -1;
    int min_b[3], div_b[3];
    for (int c = 0; c < 3; ++c) {
        min_b[c] = (int) floorf(mn[c] * inv[c]);
        div_b[c] = (int) floorf(mx[c] * inv[c]) - min_b[c] + 1;
    }
    const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
    struct Entry {
        unsigned idx;
        unsigned pt;
        bool operator<(const Entry& o) const { return idx < o.idx; }
    };
    std::vector<Entry> ev((size_t) U);
    for (long i = 0; i < U; ++i) {
        int ijk[3];
        for (int c = 0; c < 3; ++c) ijk[c] = (int) (floorf(pts[3 * i + c] * inv[c]) - (float) min_b[c]);
        ev[(size_t) i] = Entry{(unsigned) (ijk[0] * mul[0] + ijk[1] * mul[1] + ijk[2] * mul[2]), (unsigned) i};
    }
    if (order_mode == 1)
        std::sort(ev.begin(), ev.end());
    else
        std::stable_sort(ev.begin(), ev.end());
    long M = 0;
    for (size_t a = 0; a < ev.size();) {
        size_t b = a;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        while (b < ev.size() && ev[b].idx == ev[a].idx) {
            sx += pts[3 * (size_t) ev[b].pt];
            sy += pts[3 * (size_t) ev[b].pt + 1];
            sz += pts[3 * (size_t) ev[b].pt + 2];
            ++b;
        }
        const float n = (float) (b - a);
        out[3 * M] = sx / n; out[3 * M + 1] = sy / n; out[3 * M + 2] = sz / n;
        ++M;
        a = b;
    }
    return M;
}

/* Warpfield::update (:64-95): unsupported vertices -> 5 cm voxel-grid decimation -> one new node per centroid with
 * dg_se3 = calcDQB(centroid) against the OLD node set (the KD-tree is rebuilt only after the loop) and
 * dg_w = 2 * epsilon; new nodes are appended.  Output arrays need room for N + P nodes; returns the new node count. */
long orc_warpfield_update(const float* pos, const float* dq, const float* dg_w, int N, float epsilon, const float* verts,
                          long P, int blend_mode, float* pos_out, float* dq_out, float* w_out) {
    std::vector<uint8_t> flags((size_t) std::max<long>(P, 1));
    orc_unsupported(pos, dg_w, N, verts, P, flags.data());
    std::vector<float> uns;
    for (long i = 0; i < P; ++i)
        if (flags[(size_t) i]) uns.insert(uns.end(), verts + 3 * i, verts + 3 * i + 3);
    const long U = (long) uns.size() / 3;
    std::vector<float> cent((size_t) std::max<long>(3 * U, 3));
    const float leaf[3] = {(float) 0.05, 0.05f, 0.05f};
    long M = orc_voxel_grid(uns.data(), U, leaf, cent.data(), 0);
    if (M < 0) {  // PCL hands the input back unfiltered
        M = U;
        cent = uns;
    }
    std::memcpy(pos_out, pos, (size_t) N * 3 * sizeof(float));
    std::memcpy(dq_out, dq, (size_t) N * 8 * sizeof(float));
    std::memcpy(w_out, dg_w, (size_t) N * sizeof(float));
    if (M > 0) {
        std::memcpy(pos_out + 3 * (size_t) N, cent.data(), (size_t) M * 3 * sizeof(float));
        orc_blend(pos, dq, dg_w, N, cent.data(), M, blend_mode, dq_out + 8 * (size_t) N);
        for (long i = 0; i < M; ++i) w_out[N + i] = 2 * epsilon;
    }
    return N + M;
}

/* ---- TsdfVolume::raycast (src/kfusion/tsdf_volume.cpp:95-129 -> src/kfusion/cuda/tsdf_volume.cu:126-386) ---------
 * Canonical arithmetic: every operation of the reference's expressions rounded once in the order written, IEEE division
 * and sqrt, normalized(v) = v / sqrt(dot), dot = (ax bx + ay by) + az bz; fetch_tsdf clamps its index into the volume.
 * vol: packed voxels (half bits | weight << 16), x fastest.  cam2vol = {R row-major, t}; rinv = inverse rotation.
 * points4 / normals4: rows*cols float4 (NaN where the ray finds no surface); depth_or_null: u16 millimetres. */
namespace {
struct RayVol {
    const uint32_t* vol;
    int dx, dy, dz;
    float ivx, ivy, ivz;
    float at(int x, int y, int z) const { return half2float((uint16_t) (vol[(size_t) x + (size_t) y * dx + (size_t) z * dx * dy] & 0xffffu)); }
    float fetch(float px, float py, float pz) const {
        int x = (int) nearbyintf(px * ivx), y = (int) nearbyintf(py * ivy), z = (int) nearbyintf(pz * ivz);
        x = std::min(std::max(x, 0), dx - 1); y = std::min(std::max(y, 0), dy - 1); z = std::min(std::max(z, 0), dz - 1);
        return at(x, y, z);
    }
    float interp(float cx, float cy, float cz) const {
        const int gx = (int) floorf(cx), gy = (int) floorf(cy), gz = (int) floorf(cz);
        if (gx < 0 || gx >= dx - 1 || gy < 0 || gy >= dy - 1 || gz < 0 || gz >= dz - 1) return std::numeric_limits<float>::quiet_NaN();
        const float a = cx - (float) gx, b = cy - (float) gy, c = cz - (float) gz;
        const float na = 1.f - a, nb = 1.f - b, nc = 1.f - c;
        float t = 0.f;
        t += at(gx, gy, gz) * na * nb * nc;
        t += at(gx, gy, gz + 1) * na * nb * c;
        t += at(gx, gy + 1, gz) * na * b * nc;
        t += at(gx, gy + 1, gz + 1) * na * b * c;
        t += at(gx + 1, gy, gz) * a * nb * nc;
        t += at(gx + 1, gy, gz + 1) * a * nb * c;
        t += at(gx + 1, gy + 1, gz) * a * b * nc;
        t += at(gx + 1, gy + 1, gz + 1) * a * b * c;
        return t;
    }
    float interp_m(float px, float py, float pz) const { return interp(px * ivx, py * ivy, pz * ivz); }
};
inline float rdot3(float ax, float ay, float az, float bx, float by, float bz) { return (ax * bx + ay * by) + az * bz; }
}  // namespace

void orc_raycast(const uint32_t* vol, const int dims[3], const float voxel[3], float trunc, const float cam2vol[12],
                 const float rinv[9], const float intr[4], int rows, int cols, float step_factor, float grad_factor,
                 float* points4, float* normals4, uint16_t* depth_or_null) {
    RayVol V{vol, dims[0], dims[1], dims[2], 1.f / voxel[0], 1.f / voxel[1], 1.f / voxel[2]};
    const float bmx = voxel[0] * (float) dims[0] - voxel[0], bmy = voxel[1] * (float) dims[1] - voxel[1],
                bmz = voxel[2] * (float) dims[2] - voxel[2];
    const float time_step = trunc * step_factor;
    const float gdx = voxel[0] * grad_factor, gdy = voxel[1] * grad_factor, gdz = voxel[2] * grad_factor;
    const float* R = cam2vol;
    const float ox = cam2vol[9], oy = cam2vol[10], oz = cam2vol[11];
    const float finvx = 1.f / intr[0], finvy = 1.f / intr[1], cx = intr[2], cy = intr[3];
    const float qnan = std::numeric_limits<float>::quiet_NaN();
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            float* P = points4 + 4 * ((size_t) y * cols + x);
            float* Nn = normals4 + 4 * ((size_t) y * cols + x);
            for (int c = 0; c < 4; ++c) P[c] = Nn[c] = qnan;
            if (depth_or_null) depth_or_null[(size_t) y * cols + x] = 0;
            const float ux = 1.f * ((float) x - cx) * finvx, uy = 1.f * ((float) y - cy) * finvy, uz = 1.f;
            float rdx = rdot3(R[0], R[1], R[2], ux, uy, uz), rdy = rdot3(R[3], R[4], R[5], ux, uy, uz), rdz = rdot3(R[6], R[7], R[8], ux, uy, uz);
            const float len = sqrtf(rdot3(rdx, rdy, rdz, rdx, rdy, rdz));
            rdx = rdx / len; rdy = rdy / len; rdz = rdz / len;
            const float ix = 1.f / rdx, iy = 1.f / rdy, iz = 1.f / rdz;
            const float tbx = ix * (0.f - ox), tby = iy * (0.f - oy), tbz = iz * (0.f - oz);
            const float ttx = ix * (bmx - ox), tty = iy * (bmy - oy), ttz = iz * (bmz - oz);
            const float mnx = fminf(ttx, tbx), mny = fminf(tty, tby), mnz = fminf(ttz, tbz);
            const float mxx = fmaxf(ttx, tbx), mxy = fmaxf(tty, tby), mxz = fmaxf(ttz, tbz);
            float tmin = fmaxf(fmaxf(mnx, mny), fmaxf(mnx, mnz));
            float tmax = fminf(fminf(mxx, mxy), fminf(mxx, mxz));
            tmin = fmaxf(0.f, tmin);
            if (tmin >= tmax) continue;
            tmax = tmax - time_step;
            const float vsx = rdx * time_step, vsy = rdy * time_step, vsz = rdz * time_step;
            float nx = ox + rdx * tmin, ny = oy + rdy * tmin, nz = oz + rdz * tmin;
            float tsdf_next = V.fetch(nx, ny, nz);
            for (float tcurr = tmin; tcurr < tmax; tcurr += time_step) {
                const float tsdf_curr = tsdf_next;
                const float cxm = nx, cym = ny, czm = nz;
                nx += vsx; ny += vsy; nz += vsz;
                tsdf_next = V.fetch(nx, ny, nz);
                if (tsdf_curr < 0.f && tsdf_next > 0.f) break;
                if (tsdf_curr > 0.f && tsdf_next < 0.f) {
                    const float Ft = V.interp_m(cxm, cym, czm), Ftdt = V.interp_m(nx, ny, nz);
                    const float Ts = tcurr - (time_step * Ft) / (Ftdt - Ft);
                    const float vx = ox + rdx * Ts, vy = oy + rdy * Ts, vz = oz + rdz * Ts;
                    float gx = (V.interp_m(vx + gdx, vy, vz) - V.interp_m(vx - gdx, vy, vz)) / gdx;
                    float gy = (V.interp_m(vx, vy + gdy, vz) - V.interp_m(vx, vy - gdy, vz)) / gdy;
                    float gz = (V.interp_m(vx, vy, vz + gdz) - V.interp_m(vx, vy, vz - gdz)) / gdz;
                    const float gl = sqrtf(rdot3(gx, gy, gz, gx, gy, gz));
                    gx = gx / gl; gy = gy / gl; gz = gz / gl;
                    const float prod = gx * gy * gz;
                    if (prod == prod) {
                        const float dxv = vx - ox, dyv = vy - oy, dzv = vz - oz;
                        Nn[0] = rdot3(rinv[0], rinv[1], rinv[2], gx, gy, gz);
                        Nn[1] = rdot3(rinv[3], rinv[4], rinv[5], gx, gy, gz);
                        Nn[2] = rdot3(rinv[6], rinv[7], rinv[8], gx, gy, gz);
                        Nn[3] = 0.f;
                        P[0] = rdot3(rinv[0], rinv[1], rinv[2], dxv, dyv, dzv);
                        P[1] = rdot3(rinv[3], rinv[4], rinv[5], dxv, dyv, dzv);
                        P[2] = rdot3(rinv[6], rinv[7], rinv[8], dxv, dyv, dzv);
                        P[3] = 0.f;
                        if (depth_or_null) {
                            const float mm = P[2] * 1000.f;
                            depth_or_null[(size_t) y * cols + x] = (uint16_t) std::min(std::max((int) mm, 0), 65535);
                        }
                    }
                    break;
                }
            }
        }
}

uint16_t orc_float2half(float f) { return float2half(f); }
float orc_half2float(uint16_t h) { return half2float(h); }

/* src/kfusion/cuda/tsdf_volume.cu:11-22 : pack_tsdf(0.f, 0) == 0x00000000 */
void orc_tsdf_clear(uint32_t* vol, const int dims[3], int z0, int z1) {
    size_t plane = (size_t) dims[0] * dims[1];
    std::memset(vol + plane * (size_t) z0, 0, plane * (size_t) (z1 - z0) * sizeof(uint32_t));
}

/* src/kfusion/tsdf_volume.cpp:57-61 */
float orc_trunc_dist(float requested, const float voxel[3]) {
    float max_coeff = std::max(std::max(voxel[0], voxel[1]), voxel[2]);
    return std::max(requested, 2.1f * max_coeff);
}

/*
 * src/kfusion/cuda/tsdf_volume.cu:43-94 (TsdfIntegrator) + include/kfusion/cuda/device.hpp:40-45,59-67,
 * with the warp of src/dynfu/warp_field.cpp:127-148 + dual_quaternion.hpp:204-215 inserted before vol2cam.
 *
 * CANONICAL ARITHMETIC: the reference kernel accumulates vc += zstep and uses __fdividef;
 * neither is reproducible on a CPU nor invariant under z-slab sharding, and no reference test pins them.
 * Here vc is computed directly per voxel with a fixed fmaf chain, IEEE division and sqrt.
 */
long orc_tsdf_integrate(uint32_t* vol, const int dims[3], const float voxel[3], float trunc, int max_weight,
                        const float vol2cam[12], const float intr[4], const uint16_t* dists, size_t pitch_bytes,
                        int rows, int cols, const float* pos, const float* dq, const float* dg_w, int N,
                        int blend_mode, int z0, int z1, float* f32_out) {
    std::unique_ptr<KnnIndex> index;
    if (pos && N > 0) index.reset(new KnnIndex(pos, N));
    const float* R = vol2cam;          /* row-major 3x3 */
    const float* T = vol2cam + 9;      /* translation   */
    const float fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
    const float trunc_inv = 1.f / trunc; /* tsdf_volume.cu:106 */
    const size_t plane = (size_t) dims[0] * dims[1];
    long touched = 0;
#pragma omp parallel for collapse(2) schedule(dynamic, 4) reduction(+ : touched)
    for (int z = z0; z < z1; ++z)
        for (int y = 0; y < dims[1]; ++y)
            for (int x = 0; x < dims[0]; ++x) {
                float p[3] = {(float) x * voxel[0], (float) y * voxel[1], (float) z * voxel[2]};
                if (index) {
                    DQ b = blend_point(*index, pos, dq, dg_w, p, blend_mode);
                    V3 w = dq_transform_vertex(b, V3{p[0], p[1], p[2]});
                    p[0] = w.x; p[1] = w.y; p[2] = w.z;
                }
                float vcx = fmaf(R[2], p[2], fmaf(R[1], p[1], fmaf(R[0], p[0], T[0])));
                float vcy = fmaf(R[5], p[2], fmaf(R[4], p[1], fmaf(R[3], p[0], T[1])));
                float vcz = fmaf(R[8], p[2], fmaf(R[7], p[1], fmaf(R[6], p[0], T[2])));
                if (!(vcz > 0.f)) continue; /* tsdf_volume.cu:74 (vc.z <= 0) */
                float u = fmaf(fx, vcx / vcz, cx); /* device.hpp:40-45 */
                float v = fmaf(fy, vcy / vcz, cy);
                if (!(u >= 0.f && v >= 0.f && u < (float) cols && v < (float) rows)) continue; /* :70 */
                int ui = (int) u, vi = (int) v; /* point-filtered tex2D: texel floor(coo) */
                uint16_t hd = *(const uint16_t*) ((const char*) dists + (size_t) vi * pitch_bytes + (size_t) ui * 2);
                float Dp = half2float(hd);
                if (Dp == 0.f) continue; /* :74 */
                float sdf = Dp - sqrtf(fmaf(vcz, vcz, fmaf(vcy, vcy, vcx * vcx))); /* :77 */
                if (sdf >= -trunc) {                                               /* :79 */
                    float tsdf = fminf(1.f, sdf * trunc_inv);                      /* :80 */
                    size_t vi_lin = (size_t) x + (size_t) y * dims[0] + plane * (size_t) z;
                    uint32_t packed = vol[vi_lin];
                    float tsdf_prev = half2float((uint16_t) (packed & 0xffffu)); /* ushort2.x = half bits */
                    int weight_prev = (int) (packed >> 16);                      /* ushort2.y = weight    */
                    float tsdf_new = fmaf(tsdf_prev, (float) weight_prev, tsdf) / (float) (weight_prev + 1); /* :86 */
                    int weight_new = std::min(weight_prev + 1, max_weight);                                 /* :87 */
                    vol[vi_lin] = (uint32_t) float2half(tsdf_new) | ((uint32_t) weight_new << 16);
                    if (f32_out) {
                        f32_out[2 * vi_lin] = tsdf_new;
                        f32_out[2 * vi_lin + 1] = (float) weight_new;
                    }
                    ++touched;
                }
            }
    return touched;
}

float orc_tukey(float tukey_offset, float c, const float err[3]) { return tukey(tukey_offset, c, err); }
/* opt_solver.cpp:233-239 */
float orc_huber(float k, float e) {
    if (std::fabs(e) <= k) return 1.f;
    return k / std::fabs(e);
}

/* opt_solver.cpp:241-268 : the inner loop overwrites h[i] for every neighbour, the last one survives */
void orc_huber_weights(const float* pos, const float* dq, int N, float psi_reg, float* out) {
    KnnIndex index(pos, N);
    for (int i = 0; i < N; ++i) {
        int32_t nb[KNN];
        int n = index.query(pos + 3 * (size_t) i, KNN, nb, nullptr);
        float h = 0.f;
        for (int k = 0; k < n; ++k) {
            const int j = nb[k];
            V3 c{pos[3 * (size_t) j], pos[3 * (size_t) j + 1], pos[3 * (size_t) j + 2]};
            V3 a = dq_transform_vertex(load_dq(dq + 8 * (size_t) i), c);
            V3 b = dq_transform_vertex(load_dq(dq + 8 * (size_t) j), c);
            float ex = a.x - b.x, ey = a.y - b.y, ez = a.z - b.z;
            float e = sqrtf(ex * ex + ey * ey + ez * ez);
            h = orc_huber(psi_reg, e);
        }
        out[i] = h;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Solver: energy.t restated as a sparse linear least-squares problem, double precision        */

namespace {
struct Problem {
    int N;
    long P;
    std::vector<int32_t> nbr;   /* P*8  data graph, opt_solver.cpp:56-72  */
    std::vector<double> w;      /* P*8  w(canon[v], n_k), energy.t:15-17,49-52 */
    std::vector<int32_t> nnbr;  /* N*8  reg graph,  opt_solver.cpp:74-105 */
    std::vector<double> d;      /* P*3  live - canon */
    std::vector<double> theta;  /* P    tukey */
    double wreg2;               /* w_reg^2 = lambda/(N*8), opt_solver.cpp:30 */
    /* transposed data graph (node -> (point,k)) for a race-free, deterministic J^T product */
    std::vector<long> tptr;
    std::vector<long> tent;
    /* transposed reg graph (node m -> nodes n that list m) */
    std::vector<long> rptr;
    std::vector<int32_t> rent;
    const orc_solver_params* prm;
};

void build_problem(Problem& pb, const float* pos, const float* dg_w, int N, const float* canon, const float* live,
                   long P, const orc_solver_params* prm) {
    pb.N = N;
    pb.P = P;
    pb.prm = prm;
    pb.nbr.assign((size_t) P * KNN, 0);
    pb.w.assign((size_t) P * KNN, 0.0);
    pb.nnbr.assign((size_t) N * KNN, 0);
    pb.d.assign((size_t) P * 3, 0.0);
    pb.theta.assign((size_t) P, 1.0);
    pb.wreg2 = (double) prm->lambda / ((double) N * KNN);
    KnnIndex index(pos, N);
#pragma omp parallel for schedule(static)
    for (long v = 0; v < P; ++v) {
        int32_t nb[KNN];
        int n = index.query(canon + 3 * v, KNN, nb, nullptr);
        for (int k = 0; k < KNN; ++k) {
            int j = k < n ? nb[k] : nb[0];
            pb.nbr[v * KNN + k] = j;
            pb.w[v * KNN + k] = k < n ? (double) node_weight(pos + 3 * (size_t) j, dg_w[j], canon + 3 * v) : 0.0;
        }
        for (int c = 0; c < 3; ++c) pb.d[v * 3 + c] = (double) live[v * 3 + c] - (double) canon[v * 3 + c];
    }
#pragma omp parallel for schedule(static)
    for (int a = 0; a < N; ++a) {
        int32_t nb[KNN];
        int n = index.query(pos + 3 * (size_t) a, KNN, nb, nullptr);
        for (int k = 0; k < KNN; ++k) pb.nnbr[(size_t) a * KNN + k] = k < n ? nb[k] : a;
    }
    /* transposes by counting sort */
    pb.tptr.assign((size_t) N + 1, 0);
    for (long e = 0; e < P * KNN; ++e) pb.tptr[pb.nbr[e] + 1]++;
    for (int a = 0; a < N; ++a) pb.tptr[a + 1] += pb.tptr[a];
    pb.tent.assign((size_t) P * KNN, 0);
    {
        std::vector<long> cur(pb.tptr.begin(), pb.tptr.end() - 1);
        for (long e = 0; e < P * KNN; ++e) pb.tent[cur[pb.nbr[e]]++] = e;
    }
    pb.rptr.assign((size_t) N + 1, 0);
    for (long e = 0; e < (long) N * KNN; ++e) pb.rptr[pb.nnbr[e] + 1]++;
    for (int a = 0; a < N; ++a) pb.rptr[a + 1] += pb.rptr[a];
    pb.rent.assign((size_t) N * KNN, 0);
    {
        std::vector<long> cur(pb.rptr.begin(), pb.rptr.end() - 1);
        for (long e = 0; e < (long) N * KNN; ++e) pb.rent[cur[pb.nnbr[e]]++] = (int32_t) (e / KNN);
    }
}

/* residual of energy.t:47-55 without the sqrt(tukey) factor: e_v = live - canon - sum_k w_k t[n_k] */
inline void point_residual(const Problem& pb, const double* t, long v, double e[3]) {
    double s[3] = {0, 0, 0};
    for (int k = 0; k < KNN; ++k) {
        const double wk = pb.w[v * KNN + k];
        const double* tk = t + 3 * (size_t) pb.nbr[v * KNN + k];
        s[0] += wk * tk[0]; s[1] += wk * tk[1]; s[2] += wk * tk[2];
    }
    for (int c = 0; c < 3; ++c) e[c] = pb.d[v * 3 + c] - s[c];
}

void update_tukey(Problem& pb, const double* t) {
#pragma omp parallel for schedule(static)
    for (long v = 0; v < pb.P; ++v) {
        double e[3];
        point_residual(pb, t, v, e);
        float ef[3] = {(float) e[0], (float) e[1], (float) e[2]};
        pb.theta[v] = (double) tukey(pb.prm->tukey_offset, pb.prm->psi_data, ef);
    }
}

double energy(const Problem& pb, const double* t) {
    double E = 0;
#pragma omp parallel for schedule(static) reduction(+ : E)
    for (long v = 0; v < pb.P; ++v) {
        double e[3];
        point_residual(pb, t, v, e);
        E += pb.theta[v] * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    }
    double Er = 0;
    if (pb.wreg2 > 0) {
#pragma omp parallel for schedule(static) reduction(+ : Er)
        for (int a = 0; a < pb.N; ++a)
            for (int k = 0; k < KNN; ++k) { /* energy.t:73-78 : w_reg * (t[v_i] - t[n]) */
                int m = pb.nnbr[(size_t) a * KNN + k];
                for (int c = 0; c < 3; ++c) {
                    double df = t[3 * (size_t) m + c] - t[3 * (size_t) a + c];
                    Er += pb.wreg2 * df * df;
                }
            }
    }
    return E + Er;
}

/* y = (W^T Theta W + wreg2 * L) x, x and y are N*3 */
void apply_A(const Problem& pb, const double* x, double* y, std::vector<double>& s) {
    s.resize((size_t) pb.P * 3);
#pragma omp parallel for schedule(static)
    for (long v = 0; v < pb.P; ++v) {
        double a[3] = {0, 0, 0};
        for (int k = 0; k < KNN; ++k) {
            const double wk = pb.w[v * KNN + k];
            const double* xk = x + 3 * (size_t) pb.nbr[v * KNN + k];
            a[0] += wk * xk[0]; a[1] += wk * xk[1]; a[2] += wk * xk[2];
        }
        for (int c = 0; c < 3; ++c) s[v * 3 + c] = pb.theta[v] * a[c];
    }
#pragma omp parallel for schedule(static)
    for (int a = 0; a < pb.N; ++a) {
        double acc[3] = {0, 0, 0};
        for (long q = pb.tptr[a]; q < pb.tptr[a + 1]; ++q) {
            long e = pb.tent[q];
            long v = e / KNN;
            for (int c = 0; c < 3; ++c) acc[c] += pb.w[e] * s[v * 3 + c];
        }
        if (pb.wreg2 > 0) {
            for (int k = 0; k < KNN; ++k) { /* edges (a -> m): d/dt_a of (t_m - t_a)^2 */
                int m = pb.nnbr[(size_t) a * KNN + k];
                for (int c = 0; c < 3; ++c) acc[c] += pb.wreg2 * (x[3 * (size_t) a + c] - x[3 * (size_t) m + c]);
            }
            for (long q = pb.rptr[a]; q < pb.rptr[a + 1]; ++q) { /* edges (n -> a) */
                int n = pb.rent[q];
                for (int c = 0; c < 3; ++c) acc[c] += pb.wreg2 * (x[3 * (size_t) a + c] - x[3 * (size_t) n + c]);
            }
        }
        for (int c = 0; c < 3; ++c) y[3 * (size_t) a + c] = acc[c];
    }
}

void diag_A(const Problem& pb, double* dg) {
#pragma omp parallel for schedule(static)
    for (int a = 0; a < pb.N; ++a) {
        double acc = 0;
        for (long q = pb.tptr[a]; q < pb.tptr[a + 1]; ++q) {
            long e = pb.tent[q];
            acc += pb.theta[e / KNN] * pb.w[e] * pb.w[e];
        }
        if (pb.wreg2 > 0) {
            for (int k = 0; k < KNN; ++k)
                if (pb.nnbr[(size_t) a * KNN + k] != a) acc += pb.wreg2;
            for (long q = pb.rptr[a]; q < pb.rptr[a + 1]; ++q)
                if (pb.rent[q] != a) acc += pb.wreg2;
        }
        dg[a] = acc;
    }
}

double dot3(const double* a, const double* b, size_t n) {
    double s = 0;
    for (size_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}
} /* namespace */

int orc_solve(const float* pos, float* dq_inout, const float* dg_w, int N, const float* canon, const float* live,
              long P, const orc_solver_params* prm, double* t_out, double* stats_out) {
    if (N < KNN) return 1; /* precondition (reference UB at opt_solver.cpp:63-66) */
    Problem pb;
    build_problem(pb, pos, dg_w, N, canon, live, P, prm);
    const size_t n3 = (size_t) N * 3;
    std::vector<double> t(n3, 0.0) /* opt_solver.cpp:192-193 */, b(n3), r(n3), z(n3), p(n3), q(n3), dg((size_t) N), s, dl(n3);
    update_tukey(pb, t.data());
    double E0 = energy(pb, t.data());
    double E = E0;
    long pcg_total = 0, gn_total = 0;
    double rz_ref = -1.0; /* scale of the first GN step's preconditioned residual: the stop reference */
    for (int outer = 0; outer < prm->num_iter; ++outer) {
        update_tukey(pb, t.data()); /* preNonlinearSolve, opt_solver.cpp:135-140 */
        double E_outer = energy(pb, t.data());
        for (int gn = 0; gn < prm->nonlinear_iter; ++gn) {
            /* rhs = -J^T r = W^T Theta e(t) - reg * t  (A t includes both) */
            apply_A(pb, t.data(), q.data(), s);
            /* b0 = W^T Theta d */
#pragma omp parallel for schedule(static)
            for (int a = 0; a < N; ++a) {
                double acc[3] = {0, 0, 0};
                for (long qq = pb.tptr[a]; qq < pb.tptr[a + 1]; ++qq) {
                    long e = pb.tent[qq];
                    long v = e / KNN;
                    for (int c = 0; c < 3; ++c) acc[c] += pb.w[e] * pb.theta[v] * pb.d[v * 3 + c];
                }
                for (int c = 0; c < 3; ++c) b[3 * (size_t) a + c] = acc[c] - q[3 * (size_t) a + c];
            }
            diag_A(pb, dg.data());
            std::fill(dl.begin(), dl.end(), 0.0);
            r = b;
            for (int a = 0; a < N; ++a)
                for (int c = 0; c < 3; ++c) z[3 * (size_t) a + c] = dg[a] > 0 ? r[3 * (size_t) a + c] / dg[a] : 0.0;
            p = z;
            double rz = dot3(r.data(), z.data(), n3);
            if (rz_ref < 0) rz_ref = rz;
            const double rz_stop = prm->pcg_tol * prm->pcg_tol * rz_ref;
            for (int it = 0; it < prm->linear_iter && rz > 0; ++it) {
                if (rz <= rz_stop) break; /* converged relative to the problem's initial scale */
                apply_A(pb, p.data(), q.data(), s);
                double pq = dot3(p.data(), q.data(), n3);
                if (!(pq > 0)) break;
                double alpha = rz / pq;
                for (size_t i = 0; i < n3; ++i) {
                    dl[i] += alpha * p[i];
                    r[i] -= alpha * q[i];
                }
                for (int a = 0; a < N; ++a)
                    for (int c = 0; c < 3; ++c) z[3 * (size_t) a + c] = dg[a] > 0 ? r[3 * (size_t) a + c] / dg[a] : 0.0;
                double rz_new = dot3(r.data(), z.data(), n3);
                double beta = rz_new / rz;
                rz = rz_new;
                for (size_t i = 0; i < n3; ++i) p[i] = z[i] + beta * p[i];
                ++pcg_total;
            }
            for (size_t i = 0; i < n3; ++i) t[i] += dl[i];
            ++gn_total;
            double E_new = energy(pb, t.data());
            bool conv = std::fabs(E - E_new) <= 1e-12 * std::max(E_new, 1e-300);
            E = E_new;
            if (prm->early_out && conv) break;
        }
        if (prm->early_out && std::fabs(E_outer - E) <= 1e-12 * std::max(E, 1e-300) && outer > 0) break;
    }
    for (size_t i = 0; i < n3; ++i) t_out[i] = t[i];
    /* write back ONCE: opt_solver.cpp:270-285 + node.cpp:19-23 : dg_se3 := DQ(0,0,0,t) * dg_se3 */
    for (int a = 0; a < N; ++a) {
        DQ inc = dq_from_euler(0.f, 0.f, 0.f, (float) t[3 * (size_t) a], (float) t[3 * (size_t) a + 1],
                               (float) t[3 * (size_t) a + 2]);
        store_dq(dq_mul(inc, load_dq(dq_inout + 8 * (size_t) a)), dq_inout + 8 * (size_t) a);
    }
    if (stats_out) {
        stats_out[0] = E0; stats_out[1] = E; stats_out[2] = (double) pcg_total; stats_out[3] = (double) gn_total;
    }
    return 0;
}

double orc_energy(const float* pos, const float* dg_w, int N, const float* canon, const float* live, long P,
                  const orc_solver_params* prm, const double* t, const double* t_tukey) {
    Problem pb;
    build_problem(pb, pos, dg_w, N, canon, live, P, prm);
    update_tukey(pb, t_tukey);
    return energy(pb, t);
}

/* ======================================================================================================================
 * North-star extension P2PLANE_SE3 (BASELINE.json north_star (4)): a point-to-plane
 * data term with one rigid increment per node.  There is NO reference implementation (energy.t is translation-only,
 * point-to-point) -- parity unpinned; this double-precision Gauss-Newton / block-Jacobi PCG is the yardstick for the GPU
 * path and is itself cross-checked with scipy.optimize.least_squares in the tests.
 *
 *   X_k in SE(3) per node (starts at identity), ŵ_vk = w_vk / sum_j w_vj (the reference's Gaussian weights, normalised)
 *   p_v(X) = sum_k ŵ_vk X_k c_v                                  (linear blend of the node increments)
 *   E(X)   = sum_v theta_v (n_v . (p_v - l_v))^2  +  w_reg^2 sum_n sum_{m in nbr(n), m != n} | X_n g_m - X_m g_m |^2
 *   GN step: X_k <- exp(xi_k) X_k, xi = (omega, tau); rows  sqrt(theta) ŵ_vk [ (X_k c_v) x n_v ; n_v ]  and
 *            [ -[X_n g_m]x  I ] for node n, [ [X_m g_m]x  -I ] for node m of an edge.
 *   theta_v = calcTukeyBiweight(| p_v - l_v |) re-evaluated once per outer iteration, like the reference does for its own term.
 *   At the end the increments are composed onto the nodes ONCE: dg_se3_k := DQ(X_k) * dg_se3_k.
 * ==================================================================================================================== */
namespace {
struct SE3d {
    double R[9], t[3];
};
inline void se3_identity(SE3d& X) {
    for (int i = 0; i < 9; ++i) X.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    X.t[0] = X.t[1] = X.t[2] = 0.0;
}
inline void se3_apply(const SE3d& X, const double c[3], double o[3]) {
    for (int r = 0; r < 3; ++r) o[r] = X.R[3 * r] * c[0] + X.R[3 * r + 1] * c[1] + X.R[3 * r + 2] * c[2] + X.t[r];
}
inline void cross3(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
/* exp of a twist (omega, tau): R = exp([omega]x), t = V tau */
inline void se3_exp(const double xi[6], SE3d& X) {
    const double wx = xi[0], wy = xi[1], wz = xi[2];
    const double th2 = wx * wx + wy * wy + wz * wz, th = std::sqrt(th2);
    double A, B, C;  /* sin(th)/th, (1-cos th)/th^2, (th - sin th)/th^3 */
    if (th < 1e-6) {
        A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; C = 1.0 / 6.0 - th2 / 120.0;
    } else {
        A = std::sin(th) / th; B = (1.0 - std::cos(th)) / th2; C = (th - std::sin(th)) / (th2 * th);
    }
    const double K[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double K2[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) K2[3 * r + c] = K[3 * r] * K[c] + K[3 * r + 1] * K[3 + c] + K[3 * r + 2] * K[6 + c];
    double V[9];
    for (int i = 0; i < 9; ++i) {
        const double I = (i % 4 == 0) ? 1.0 : 0.0;
        X.R[i] = I + A * K[i] + B * K2[i];
        V[i] = I + B * K[i] + C * K2[i];
    }
    for (int r = 0; r < 3; ++r) X.t[r] = V[3 * r] * xi[3] + V[3 * r + 1] * xi[4] + V[3 * r + 2] * xi[5];
}
inline SE3d se3_mul(const SE3d& A, const SE3d& B) { /* A o B */
    SE3d C;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) C.R[3 * r + c] = A.R[3 * r] * B.R[c] + A.R[3 * r + 1] * B.R[3 + c] + A.R[3 * r + 2] * B.R[6 + c];
        C.t[r] = A.R[3 * r] * B.t[0] + A.R[3 * r + 1] * B.t[1] + A.R[3 * r + 2] * B.t[2] + A.t[r];
    }
    return C;
}
/* rotation matrix -> unit quaternion (w,x,y,z), Shepperd's method */
inline void rot_to_quat(const double R[9], double q[4]) {
    const double tr = R[0] + R[4] + R[8];
    if (tr > 0) {
        double s = std::sqrt(tr + 1.0) * 2;
        q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        double s = std::sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
        q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s;
    } else if (R[4] > R[8]) {
        double s = std::sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
        q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s;
    } else {
        double s = std::sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
        q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s;
    }
}
/* solve the 6x6 SPD system M z = r by Cholesky; returns false (z = 0) if M is not positive definite */
inline bool chol6_solve(const double M[36], const double r[6], double z[6]) {
    double L[36] = {0};
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = M[6 * i + j];
            for (int k = 0; k < j; ++k) s -= L[6 * i + k] * L[6 * j + k];
            if (i == j) {
                if (!(s > 0)) {
                    for (int k = 0; k < 6; ++k) z[k] = 0;
                    return false;
                }
                L[6 * i + i] = std::sqrt(s);
            } else {
                L[6 * i + j] = s / L[6 * j + j];
            }
        }
    double y[6];
    for (int i = 0; i < 6; ++i) {
        double s = r[i];
        for (int k = 0; k < i; ++k) s -= L[6 * i + k] * y[k];
        y[i] = s / L[6 * i + i];
    }
    for (int i = 5; i >= 0; --i) {
        double s = y[i];
        for (int k = i + 1; k < 6; ++k) s -= L[6 * k + i] * z[k];
        z[i] = s / L[6 * i + i];
    }
    return true;
}

struct P2P {
    const Problem* pb;
    const float *canon, *live, *nrm, *pos;
    std::vector<double> wn;   /* P*8 normalised weights */
    std::vector<SE3d> X;      /* N */
    std::vector<double> jac;  /* P*8*6: ŵ [q x n ; n] */
    std::vector<double> e;    /* P: n.(p - l) */
    std::vector<double> G;    /* N*8*2*3: for edge (n,i): X_n g_m and X_m g_m */
    const float* dg_w;
    std::vector<double> ew;   /* N*8: weight of edge (n,i): w_reg^2, times alpha_ij h_ij with reg_mode 1; 0 for self edges */
};
/* edge weights at the current X (reg_mode 1: the Huber weight is frozen between outer iterations, like the Tukey weights) */
void p2p_update_edge_weights(P2P& S) {
    const Problem& pb = *S.pb;
    S.ew.assign((size_t) pb.N * KNN, 0.0);
    for (int a = 0; a < pb.N; ++a)
        for (int i = 0; i < KNN; ++i) {
            const int m = pb.nnbr[(size_t) a * KNN + i];
            if (m == a) continue;
            double w = pb.wreg2;
            if (pb.prm->reg_mode == 1) {
                const double g[3] = {S.pos[3 * (size_t) m], S.pos[3 * (size_t) m + 1], S.pos[3 * (size_t) m + 2]};
                double x[3], y[3];
                se3_apply(S.X[a], g, x);
                se3_apply(S.X[m], g, y);
                const double r = std::sqrt((x[0] - y[0]) * (x[0] - y[0]) + (x[1] - y[1]) * (x[1] - y[1]) + (x[2] - y[2]) * (x[2] - y[2]));
                const double psi = pb.prm->psi_reg;
                w *= std::max((double) S.dg_w[a], (double) S.dg_w[m]) * (r <= psi ? 1.0 : psi / r);
            }
            S.ew[(size_t) a * KNN + i] = w;
        }
}
void p2p_point(const P2P& S, long v, double p[3]) {
    p[0] = p[1] = p[2] = 0;
    const double c[3] = {S.canon[3 * v], S.canon[3 * v + 1], S.canon[3 * v + 2]};
    for (int k = 0; k < KNN; ++k) {
        const double w = S.wn[v * KNN + k];
        if (w == 0) continue;
        double q[3];
        se3_apply(S.X[S.pb->nbr[v * KNN + k]], c, q);
        for (int r = 0; r < 3; ++r) p[r] += w * q[r];
    }
}
void p2p_update_tukey(P2P& S, Problem& pb) {
#pragma omp parallel for schedule(static)
    for (long v = 0; v < pb.P; ++v) {
        double p[3];
        p2p_point(S, v, p);
        float ef[3] = {(float) (S.live[3 * v] - p[0]), (float) (S.live[3 * v + 1] - p[1]), (float) (S.live[3 * v + 2] - p[2])};
        double sw = 0;
        for (int k = 0; k < KNN; ++k) sw += S.wn[v * KNN + k];
        pb.theta[v] = sw > 0 ? (double) tukey(pb.prm->tukey_offset, pb.prm->psi_data, ef) : 0.0;
    }
}
void p2p_linearise(P2P& S) {
    const Problem& pb = *S.pb;
#pragma omp parallel for schedule(static)
    for (long v = 0; v < pb.P; ++v) {
        const double c[3] = {S.canon[3 * v], S.canon[3 * v + 1], S.canon[3 * v + 2]};
        const double n[3] = {S.nrm[3 * v], S.nrm[3 * v + 1], S.nrm[3 * v + 2]};
        double p[3] = {0, 0, 0};
        for (int k = 0; k < KNN; ++k) {
            const double w = S.wn[v * KNN + k];
            double q[3], qxn[3];
            se3_apply(S.X[pb.nbr[v * KNN + k]], c, q);
            cross3(q, n, qxn);
            double* a = &S.jac[(v * KNN + k) * 6];
            for (int r = 0; r < 3; ++r) {
                p[r] += w * q[r];
                a[r] = w * qxn[r];
                a[3 + r] = w * n[r];
            }
        }
        S.e[v] = n[0] * (p[0] - S.live[3 * v]) + n[1] * (p[1] - S.live[3 * v + 1]) + n[2] * (p[2] - S.live[3 * v + 2]);
    }
#pragma omp parallel for schedule(static)
    for (int a = 0; a < pb.N; ++a)
        for (int i = 0; i < KNN; ++i) {
            const int m = pb.nnbr[(size_t) a * KNN + i];
            const double g[3] = {S.pos[3 * (size_t) m], S.pos[3 * (size_t) m + 1], S.pos[3 * (size_t) m + 2]};
            se3_apply(S.X[a], g, &S.G[((size_t) a * KNN + i) * 6]);
            se3_apply(S.X[m], g, &S.G[((size_t) a * KNN + i) * 6 + 3]);
        }
}
double p2p_energy(const P2P& S) {
    const Problem& pb = *S.pb;
    double E = 0;
#pragma omp parallel for schedule(static) reduction(+ : E)
    for (long v = 0; v < pb.P; ++v) {
        double p[3];
        p2p_point(S, v, p);
        const double e = S.nrm[3 * v] * (p[0] - S.live[3 * v]) + S.nrm[3 * v + 1] * (p[1] - S.live[3 * v + 1]) +
                         S.nrm[3 * v + 2] * (p[2] - S.live[3 * v + 2]);
        E += pb.theta[v] * e * e;
    }
    double Er = 0;
    for (int a = 0; a < pb.N; ++a)
        for (int i = 0; i < KNN; ++i) {
            const int m = pb.nnbr[(size_t) a * KNN + i];
            if (m == a) continue;
            const double g[3] = {S.pos[3 * (size_t) m], S.pos[3 * (size_t) m + 1], S.pos[3 * (size_t) m + 2]};
            double x[3], y[3];
            se3_apply(S.X[a], g, x);
            se3_apply(S.X[m], g, y);
            Er += S.ew[(size_t) a * KNN + i] * ((x[0] - y[0]) * (x[0] - y[0]) + (x[1] - y[1]) * (x[1] - y[1]) + (x[2] - y[2]) * (x[2] - y[2]));
        }
    return E + Er;
}
/* y = (J^T J) x for x in R^{6N}; with rhs != nullptr also rhs = -J^T r0 and the 6x6 diagonal blocks D */
void p2p_apply(const P2P& S, const double* x, double* y, double* rhs, double* D) {
    const Problem& pb = *S.pb;
    std::vector<double> s((size_t) pb.P);
#pragma omp parallel for schedule(static)
    for (long v = 0; v < pb.P; ++v) {
        double acc = 0;
        for (int k = 0; k < KNN; ++k) {
            const double* a = &S.jac[(v * KNN + k) * 6];
            const double* xk = x + 6 * (size_t) pb.nbr[v * KNN + k];
            for (int r = 0; r < 6; ++r) acc += a[r] * xk[r];
        }
        s[(size_t) v] = pb.theta[v] * acc;
    }
#pragma omp parallel for schedule(static)
    for (int n = 0; n < pb.N; ++n) {
        double acc[6] = {0}, b[6] = {0}, M[36] = {0};
        for (long qq = pb.tptr[n]; qq < pb.tptr[n + 1]; ++qq) {
            const long ent = pb.tent[qq], v = ent / KNN;
            const double* a = &S.jac[ent * 6];
            for (int r = 0; r < 6; ++r) acc[r] += a[r] * s[(size_t) v];
            if (rhs) {
                for (int r = 0; r < 6; ++r) {
                    b[r] -= pb.theta[v] * S.e[v] * a[r];
                    for (int c = 0; c < 6; ++c) M[6 * r + c] += pb.theta[v] * a[r] * a[c];
                }
            }
        }
        if (pb.wreg2 > 0) {
            /* edges where n is the source (out) and where n is the target (in) */
            auto edge = [&](int src, int i, bool as_source) {
                const int m = pb.nnbr[(size_t) src * KNN + i];
                if (m == src) return;
                const double we = S.ew[(size_t) src * KNN + i];
                const double* Gs = &S.G[((size_t) src * KNN + i) * 6];      /* X_src g_m */
                const double* Gm = Gs + 3;                                   /* X_m g_m   */
                const double* xs = x + 6 * (size_t) src;
                const double* xm = x + 6 * (size_t) m;
                double c1[3], c2[3], rho[3];
                cross3(xs, Gs, c1);
                cross3(xm, Gm, c2);
                for (int r = 0; r < 3; ++r) rho[r] = c1[r] + xs[3 + r] - c2[r] - xm[3 + r];
                const double* Gk = as_source ? Gs : Gm;
                const double sg = as_source ? 1.0 : -1.0;
                double gxr[3];
                cross3(Gk, rho, gxr);
                for (int r = 0; r < 3; ++r) {
                    acc[r] += we * sg * gxr[r];
                    acc[3 + r] += we * sg * rho[r];
                }
                if (rhs) {
                    double rho0[3] = {Gs[0] - Gm[0], Gs[1] - Gm[1], Gs[2] - Gm[2]}, gx0[3];
                    cross3(Gk, rho0, gx0);
                    for (int r = 0; r < 3; ++r) {
                        b[r] -= we * sg * gx0[r];
                        b[3 + r] -= we * sg * rho0[r];
                    }
                    /* J = sg * [ -[Gk]x  I ]  ->  J^T J = [ [Gk]x^T [Gk]x   [Gk]x ; -[Gk]x  I ]  (sign cancels) */
                    const double K[9] = {0, -Gk[2], Gk[1], Gk[2], 0, -Gk[0], -Gk[1], Gk[0], 0};
                    for (int r = 0; r < 3; ++r)
                        for (int c = 0; c < 3; ++c) {
                            double ktk = 0;
                            for (int k = 0; k < 3; ++k) ktk += K[3 * k + r] * K[3 * k + c];
                            M[6 * r + c] += we * ktk;
                            M[6 * r + 3 + c] += we * K[3 * r + c];       /* (-K)^T = K */
                            M[6 * (3 + r) + c] += we * (-K[3 * r + c]);
                            if (r == c) M[6 * (3 + r) + 3 + c] += we;
                        }
                }
            };
            for (int i = 0; i < KNN; ++i) edge(n, i, true);
            for (long qq = pb.rptr[n]; qq < pb.rptr[n + 1]; ++qq) {
                const int src = pb.rent[qq];
                for (int i = 0; i < KNN; ++i)
                    if (pb.nnbr[(size_t) src * KNN + i] == n && src != n) edge(src, i, false);
            }
        }
        for (int r = 0; r < 6; ++r) y[6 * (size_t) n + r] = acc[r];
        if (rhs) {
            for (int r = 0; r < 6; ++r) rhs[6 * (size_t) n + r] = b[r];
            for (int r = 0; r < 36; ++r) D[36 * (size_t) n + r] = M[r];
        }
    }
}
}  // namespace

int orc_solve_p2plane(const float* pos, float* dq_inout, const float* dg_w, int N, const float* canon, const float* live,
                      const float* live_n, long P, const orc_solver_params* prm, double* X_out, double* stats_out) {
    if (N < KNN) return 1;
    Problem pb;
    build_problem(pb, pos, dg_w, N, canon, live, P, prm);
    P2P S;
    S.pb = &pb; S.canon = canon; S.live = live; S.nrm = live_n; S.pos = pos; S.dg_w = dg_w;
    S.wn.assign((size_t) P * KNN, 0.0);
    for (long v = 0; v < P; ++v) {
        double sw = 0;
        for (int k = 0; k < KNN; ++k) sw += pb.w[v * KNN + k];
        for (int k = 0; k < KNN; ++k) S.wn[v * KNN + k] = sw > 0 ? pb.w[v * KNN + k] / sw : 0.0;
    }
    S.X.resize((size_t) N);
    for (auto& X : S.X) se3_identity(X);
    p2p_update_edge_weights(S);
    S.jac.assign((size_t) P * KNN * 6, 0.0);
    S.e.assign((size_t) P, 0.0);
    S.G.assign((size_t) N * KNN * 6, 0.0);
    const size_t n6 = (size_t) N * 6;
    std::vector<double> b(n6), r(n6), z(n6), p(n6), q(n6), x(n6), D((size_t) N * 36), zero(n6, 0.0);
    p2p_update_tukey(S, pb);
    const double E0 = p2p_energy(S);
    double E = E0, rz_ref = -1;
    long pcg_total = 0, gn_total = 0;
    auto precond = [&](const std::vector<double>& rr, std::vector<double>& zz) {
        for (int n = 0; n < N; ++n) chol6_solve(&D[36 * (size_t) n], &rr[6 * (size_t) n], &zz[6 * (size_t) n]);
    };
    for (int outer = 0; outer < prm->num_iter; ++outer) {
        p2p_update_tukey(S, pb);
        p2p_update_edge_weights(S);
        for (int gn = 0; gn < prm->nonlinear_iter; ++gn) {
            p2p_linearise(S);
            p2p_apply(S, zero.data(), q.data(), b.data(), D.data());
            std::fill(x.begin(), x.end(), 0.0);
            r = b;
            precond(r, z);
            p = z;
            double rz = 0;
            for (size_t i = 0; i < n6; ++i) rz += r[i] * z[i];
            if (rz_ref < 0) rz_ref = rz;
            const double rz_stop = prm->pcg_tol * prm->pcg_tol * rz_ref;
            for (int it = 0; it < prm->linear_iter && rz > 0; ++it) {
                if (rz <= rz_stop) break;
                p2p_apply(S, p.data(), q.data(), nullptr, nullptr);
                double pq = 0;
                for (size_t i = 0; i < n6; ++i) pq += p[i] * q[i];
                if (!(pq > 0)) break;
                const double alpha = rz / pq;
                for (size_t i = 0; i < n6; ++i) {
                    x[i] += alpha * p[i];
                    r[i] -= alpha * q[i];
                }
                precond(r, z);
                double rzn = 0;
                for (size_t i = 0; i < n6; ++i) rzn += r[i] * z[i];
                const double beta = rzn / rz;
                rz = rzn;
                for (size_t i = 0; i < n6; ++i) p[i] = z[i] + beta * p[i];
                ++pcg_total;
            }
            for (int n = 0; n < N; ++n) {
                SE3d dX;
                se3_exp(&x[6 * (size_t) n], dX);
                S.X[(size_t) n] = se3_mul(dX, S.X[(size_t) n]);
            }
            ++gn_total;
            E = p2p_energy(S);
        }
    }
    /* compose once onto the nodes: dg_se3 := DQ(X) * dg_se3 */
    for (int n = 0; n < N; ++n) {
        double qd[4];
        rot_to_quat(S.X[(size_t) n].R, qd);
        const float tf[3] = {(float) S.X[(size_t) n].t[0], (float) S.X[(size_t) n].t[1], (float) S.X[(size_t) n].t[2]};
        DQ inc = dq_from_rot_trans(Quat{(float) qd[0], (float) qd[1], (float) qd[2], (float) qd[3]}, tf);
        store_dq(dq_mul(inc, load_dq(dq_inout + 8 * (size_t) n)), dq_inout + 8 * (size_t) n);
        if (X_out) {
            for (int i = 0; i < 9; ++i) X_out[12 * (size_t) n + i] = S.X[(size_t) n].R[i];
            for (int i = 0; i < 3; ++i) X_out[12 * (size_t) n + 9 + i] = S.X[(size_t) n].t[i];
        }
    }
    if (stats_out) {
        stats_out[0] = E0; stats_out[1] = E; stats_out[2] = (double) pcg_total; stats_out[3] = (double) gn_total;
    }
    return 0;
}

/* the same energy for arbitrary increments X (N*12: R row-major, t), Tukey weights evaluated at X_tukey: what the tests hand
 * to scipy.optimize.least_squares */
double orc_energy_p2plane(const float* pos, const float* dg_w, int N, const float* canon, const float* live, const float* live_n,
                          long P, const orc_solver_params* prm, const double* X, const double* X_tukey) {
    Problem pb;
    build_problem(pb, pos, dg_w, N, canon, live, P, prm);
    P2P S;
    S.pb = &pb; S.canon = canon; S.live = live; S.nrm = live_n; S.pos = pos; S.dg_w = dg_w;
    S.wn.assign((size_t) P * KNN, 0.0);
    for (long v = 0; v < P; ++v) {
        double sw = 0;
        for (int k = 0; k < KNN; ++k) sw += pb.w[v * KNN + k];
        for (int k = 0; k < KNN; ++k) S.wn[v * KNN + k] = sw > 0 ? pb.w[v * KNN + k] / sw : 0.0;
    }
    auto load = [&](const double* src) {
        S.X.resize((size_t) N);
        for (int n = 0; n < N; ++n) {
            for (int i = 0; i < 9; ++i) S.X[(size_t) n].R[i] = src[12 * (size_t) n + i];
            for (int i = 0; i < 3; ++i) S.X[(size_t) n].t[i] = src[12 * (size_t) n + 9 + i];
        }
    };
    load(X_tukey);
    p2p_update_tukey(S, pb);
    p2p_update_edge_weights(S);
    load(X);
    return p2p_energy(S);
}

/* ---- marching cubes (src/kfusion/cuda/marching_cubes.cu:33-75 cube index, :77-140 occupied voxels, :181-199 cell
 * centres + edge interpolation, :201-260 triangle emission; host flow src/kfusion/marching_cubes.cpp:20-63).
 * Differences from the reference, all deliberate and stated in DESIGN.md: any volume size (the reference hard-codes 128^3,
 * internal.hpp:74, marching_cubes.cu:151-152,283-285); cubes are emitted in a FIXED order -- tiles of 32 x 8 x 8 cubes in
 * ascending (z, y, x) tile order, cubes in ascending (z, y, x) order inside a tile (the reference's order is decided by
 * atomics, :108); IEEE arithmetic without contraction.  tri: the 256 x 16 triangle table.  Returns the number of vertices that
 * exist; writes at most `capacity` of them as (x, y, z, 1) and, when cube_ids is non-NULL, the linear cube index of each. */
long orc_marching_cubes(const uint32_t* vol, const int dims[3], const float volume_size[3], const signed char* tri, float* verts4,
                        int32_t* cube_ids, long capacity) {
    const int dx = dims[0], dy = dims[1], dz = dims[2];
    const float cell[3] = {volume_size[0] / (float) dx, volume_size[1] / (float) dy, volume_size[2] / (float) dz};
    auto vox = [&](int x, int y, int z, float& f, int& w) {
        const uint32_t v = vol[(size_t) x + (size_t) dx * ((size_t) y + (size_t) dy * (size_t) z)];
        f = half2float((uint16_t) (v & 0xffffu));
        w = (int) (v >> 16);
    };
    static const int CX[8] = {0, 1, 1, 0, 0, 1, 1, 0}, CY[8] = {0, 0, 1, 1, 0, 0, 1, 1}, CZ[8] = {0, 0, 0, 0, 1, 1, 1, 1};
    static const int EA[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3}, EB[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};
    long n = 0;
    const int ntx = (dx - 1 + 31) / 32, nty = (dy - 1 + 7) / 8, ntz = (dz - 1 + 7) / 8;
    for (int tz = 0; tz < ntz; ++tz)
        for (int ty = 0; ty < nty; ++ty)
            for (int tx = 0; tx < ntx; ++tx)
                for (int z = tz * 8; z < std::min(tz * 8 + 8, dz - 1); ++z)
                    for (int y = ty * 8; y < std::min(ty * 8 + 8, dy - 1); ++y)
                        for (int x = tx * 32; x < std::min(tx * 32 + 32, dx - 1); ++x) {
                            float f[8];
                            int cubeindex = 0;
                            bool ok = true;
                            for (int c = 0; c < 8 && ok; ++c) {  /* computeCubeIndex: 0 at the first corner without weight */
                                int w;
                                vox(x + CX[c], y + CY[c], z + CZ[c], f[c], w);
                                if (w == 0) ok = false;
                            }
                            if (!ok) continue;
                            for (int c = 0; c < 8; ++c) cubeindex += (f[c] < 0.f) ? (1 << c) : 0;
                            const signed char* row = tri + 16 * cubeindex;
                            if (row[0] < 0) continue;
                            float p[8][3];
                            for (int c = 0; c < 8; ++c) { /* getNodeCoo: (i + 0.5) * cell */
                                p[c][0] = ((float) (x + CX[c]) + 0.5f) * cell[0];
                                p[c][1] = ((float) (y + CY[c]) + 0.5f) * cell[1];
                                p[c][2] = ((float) (z + CZ[c]) + 0.5f) * cell[2];
                            }
                            for (int i = 0; i < 16 && row[i] >= 0; ++i) {
                                const int e = row[i], a = EA[e], b = EB[e];
                                /* vertex_interp: t = (iso - f0) / (f1 - f0 + 1e-15f), p0 + t (p1 - p0) */
                                const float t = (0.f - f[a]) / ((f[b] - f[a]) + 1e-15f);
                                if (n < capacity) {
                                    for (int k = 0; k < 3; ++k) verts4[4 * n + k] = p[a][k] + t * (p[b][k] - p[a][k]);
                                    verts4[4 * n + 3] = 1.f;
                                    if (cube_ids) cube_ids[n] = (int32_t) ((size_t) x + (size_t) dx * ((size_t) y + (size_t) dy * (size_t) z));
                                }
                                ++n;
                            }
                        }
    return n;
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm of the bench asks for all host threads explicitly */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void) n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

} /* extern "C" */
