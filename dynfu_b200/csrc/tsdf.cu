// TSDF volume kernels: compute_dists, clear, rigid and WARPED projective integration.
//
// Replaces kfusion::device::{compute_dists, clear_volume, integrate}
// (src/kfusion/cuda/imgproc.cu:233-254, src/kfusion/cuda/tsdf_volume.cu:11-34, 43-121) and inserts the
// per-voxel warp Warpfield::calcDQB(p).transformVertex(p) (src/dynfu/warp_field.cpp:127-148,
// include/dynfu/utils/dual_quaternion.hpp:204-215) in front of vol2cam.
//
// Layout in HBM: the reference's own ushort2{half tsdf, u16 weight} volume, x fastest.  One CTA owns a
// 32x8x8-voxel tile = four 8x8x8 bricks.  Every thread owns quads of 4 x-consecutive voxels, so volume
// traffic is 128-bit ld/st (streaming, evict-first) and a warp of the rigid pass covers whole 128-byte
// lines.  A brick is "near" when some voxel of it may have a non-zero node weight; near bricks run the
// exact per-voxel 8-NN + blend, all other bricks are provably un-warped and take the rigid path.
#include <algorithm>

#include "blend.cuh"
#include "dfu_internal.h"

using namespace dfu;

namespace {

struct IntegrateArgs {
    uint32_t* vol;
    int dx, dy, dz;
    float vsx, vsy, vsz;
    float trunc, trunc_inv;
    int max_weight;
    float R[9], T[3];
    float fx, fy, cx, cy;
    const uint16_t* dists;
    size_t pitch;
    int rows, cols;
    int z0, z1, zt0;  // planes [z0,z1); zt0 = first tile plane (multiple of 8)
    int ntx, nty;
    // warp field (warped != 0)
    int warped, blend_mode;
    const float4 *pos_w, *real, *dual;
    int N;
    const int* flags;
    const float2* bounds;
    int bdx, bdy;
    uint4* knn_pool;       // per-voxel 8-NN id cache (null: disabled), see BrickTable; indexed by GLOBAL brick id
    float4* w_pool;        // per-voxel weight cache (two float4 per voxel)
    unsigned char* built;
    // per-call scratch written by tile_classify_kernel
    unsigned short* tile_flags; // per 32x8x8 tile: bits 0-3 = near mask of its 4 bricks, bit 4 = rigid pass cannot touch a voxel,
                                // bits 8-11 = bricks all of whose voxels are PROVABLY updated with tsdf = 1 (free space in front of the surface)
    int* work_count;            // [0] tiles with any work, [1] tiles with a near brick missing from the 8-NN cache,
                                // [2],[3] ticket counters of the two lists
    int* work_tiles;            // ids of the tiles with work
    int* fill_tiles;            // ids of the tiles to fill
    int ntiles;
    const float2* dmax_tiles;  // per 16x16-pixel tile of the dists image: .x = max ray length (0: no depth in the tile),
                               // .y = min ray length (0: some pixel of the tile has no depth)
    int dtx, dty;             // tiles per row / column
    unsigned long long* stats;  // instrumentation of the last call on this device: [0] voxels updated, [1] quads (16 B) read + written,
                                // [2] voxels updated through the saturated-free-space path, [3] bricks that ran the per-voxel warp (zeroed by tile_classify_kernel)
};
struct Tally {
    unsigned vox = 0, quads = 0, bricks = 0, sat = 0;
};

constexpr int DT = 16;  // pixels per side of a depth tile

DFU_DEV uint4 ld_stream(const uint4* p) { return __ldcs(p); }
// TMA bulk prefetch of a contiguous block into L2 (one instruction per block, no registers, no shared memory):
// cp.async.bulk.prefetch.L2 -- bytes % 16 == 0, 16-byte aligned source
DFU_DEV void tma_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
DFU_DEV void st_stream(uint4* p, uint4 v) { __stcs(p, v); }

// TsdfIntegrator::operator() body for one voxel whose (warped) volume-frame position is (px,py,pz)
// (src/kfusion/cuda/tsdf_volume.cu:66-80, Projector include/kfusion/cuda/device.hpp:40-45) in the canonical
// arithmetic of DESIGN.md: direct fmaf chain for vc, IEEE division and sqrt.  Returns true and the
// truncated sdf when the voxel is to be updated.
DFU_DEV bool voxel_tsdf(const IntegrateArgs& a, float px, float py, float pz, float& tsdf) {
    const float vcx = __fmaf_rn(a.R[2], pz, __fmaf_rn(a.R[1], py, __fmaf_rn(a.R[0], px, a.T[0])));
    const float vcy = __fmaf_rn(a.R[5], pz, __fmaf_rn(a.R[4], py, __fmaf_rn(a.R[3], px, a.T[1])));
    const float vcz = __fmaf_rn(a.R[8], pz, __fmaf_rn(a.R[7], py, __fmaf_rn(a.R[6], px, a.T[2])));
    if (!(vcz > 0.f)) return false;  // tsdf_volume.cu:74 (vc.z <= 0)
    const float u = __fmaf_rn(a.fx, __fdiv_rn(vcx, vcz), a.cx);
    const float v = __fmaf_rn(a.fy, __fdiv_rn(vcy, vcz), a.cy);
    if (!(u >= 0.f && v >= 0.f && u < (float) a.cols && v < (float) a.rows)) return false;  // :70
    const int ui = (int) u, vi = (int) v;  // point-filtered tex2D: texel floor(coo)
    const unsigned short hd =
        __ldg(reinterpret_cast<const unsigned short*>(reinterpret_cast<const char*>(a.dists) + (size_t) vi * a.pitch) + ui);
    const float Dp = __half2float(__ushort_as_half(hd));
    if (Dp == 0.f) return false;  // :74
    const float sdf = fsub(Dp, __fsqrt_rn(__fmaf_rn(vcz, vcz, __fmaf_rn(vcy, vcy, fmul(vcx, vcx)))));  // :77
    if (!(sdf >= -a.trunc)) return false;                                                           // :79
    tsdf = fminf(1.f, fmul(sdf, a.trunc_inv));                                                      // :80
    return true;
}

// running average (tsdf_volume.cu:83-90)
DFU_DEV uint32_t voxel_update(const IntegrateArgs& a, uint32_t packed, float tsdf) {
    int weight_prev;
    const float tsdf_prev = unpack_tsdf(packed, weight_prev);
    const float tsdf_new = __fdiv_rn(__fmaf_rn(tsdf_prev, (float) weight_prev, tsdf), (float) (weight_prev + 1));
    const int weight_new = min(weight_prev + 1, a.max_weight);
    return pack_tsdf(tsdf_new, weight_new);
}

// the same update for tsdf == 1 (saturated free space).  Two cases need no arithmetic and give the same bits as the general
// expression: a voxel that already holds 1.0 stays 1.0 (fma(1, W, 1) = W + 1 exactly, (W + 1) / (W + 1) = 1), and a voxel
// without weight takes the new value ((0 * 0 + 1) / 1 = 1; a cleared voxel holds +0.0).  That is every free-space voxel after
// its first frame.
DFU_DEV uint32_t voxel_update_one(const IntegrateArgs& a, uint32_t packed) {
    const uint32_t w = packed >> 16, h = packed & 0xffffu;
    if (h == 0x3c00u || (w == 0u && h == 0u)) return 0x3c00u | ((uint32_t) min((int) w + 1, a.max_weight) << 16);
    return voxel_update(a, packed, 1.f);
}

DFU_DEV void quad_commit(const IntegrateArgs& a, size_t lin, const bool (&hit)[4], const float (&ts)[4], Tally& tally) {
    if (!(hit[0] | hit[1] | hit[2] | hit[3])) return;  // untouched quads cost no volume traffic
    tally.vox += (unsigned) hit[0] + (unsigned) hit[1] + (unsigned) hit[2] + (unsigned) hit[3];
    tally.quads += 1;
    uint4* p = reinterpret_cast<uint4*>(a.vol + lin);
    uint4 v = ld_stream(p);
    if (hit[0]) v.x = voxel_update(a, v.x, ts[0]);
    if (hit[1]) v.y = voxel_update(a, v.y, ts[1]);
    if (hit[2]) v.z = voxel_update(a, v.z, ts[2]);
    if (hit[3]) v.w = voxel_update(a, v.w, ts[3]);
    st_stream(p, v);
}

constexpr int CHUNK = 1024;          // nodes examined per candidate-compaction round
constexpr int SEG = CHUNK / 4;       // per-warp segment of the candidate buffer (4 warps)

struct IntegrateSmem {
    float4 cand[CHUNK];  // (x, y, z, index bits) of the candidate nodes of the current chunk
    int cnt[4];
};

// what one CTA needs to know about its 32x8x8 tile
struct TileInfo {
    int x0, y0, zt;
    size_t brick0;          // index of the tile's first brick in the brick table
    int near_mask;          // bit sb set: brick sb needs the exact per-voxel warp
    int sat_mask;           // bit sb set: every voxel of brick sb is updated with tsdf = 1 (no per-voxel geometry needed)
    bool translation_only;  // every node is a pure translation
    float reff2;            // squared distance beyond which rule (b) below holds for a single voxel
    float r_brick;
};

// uniform per-call scalars of the warp field (rule (b) radius etc.)
struct FieldInfo {
    bool translation_only;
    bool all_near;
    float r_zero, r_eff, reff2;
    float maxw, dmax;  // largest dg_w, largest |dual.xyz| component
};
// A brick must run the exact per-voxel warp ("near") unless every voxel of it is PROVABLY left where it is:
//  (a) all 8 weights are exactly 0.f beyond sqrt(209)*dg_w of every node (dfu_math.cuh node_weight), or
//  (b) translation-only field: p' = fl(p + 2*acc) with |2*acc_c| <= 16*dmax*w, w <= exp(-dmin^2/(2 maxw^2));
//      when that is below p_c * 2^-26 (less than half an ulp of p_c >= voxel size) the addition returns p_c
//      bit for bit.  Bricks touching index 0 of an axis (p_c == 0) are excluded from (b).
DFU_DEV FieldInfo field_info(const IntegrateArgs& a) {
    FieldInfo f{false, false, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (!a.warped) return f;
    f.translation_only = a.flags[0] != 0;
    const float maxw = __int_as_float(a.flags[1]);
    const float dmax = __int_as_float(a.flags[2]);
    f.maxw = maxw;
    f.dmax = dmax;
    f.r_zero = 14.4569f * maxw * 1.001f + 1e-6f;  // rule (a); 1.001 covers rounding
    f.all_near = (a.blend_mode == DFU_BLEND_REF_COMPOSE) && !f.translation_only;
    f.r_eff = f.r_zero;
    if (f.translation_only && a.blend_mode == DFU_BLEND_REF_COMPOSE) {
        const float vmin = fminf(a.vsx, fminf(a.vsy, a.vsz));
        // 16*dmax*exp(-x) * 1.0001 < vmin * 2^-26  <=>  x > L
        const float L = logf(fmaxf(16.f * dmax * 1.0001f, 1e-37f)) - logf(vmin * 1.4901161e-8f) + 1e-3f;
        const float re = L > 0.f ? maxw * sqrtf(2.f * L) * 1.0001f + 1e-6f : 0.f;
        f.r_eff = fminf(f.r_zero, re);
        f.reff2 = f.r_eff * f.r_eff;
    }
    return f;
}

DFU_DEV TileInfo tile_info(const IntegrateArgs& a, int tile, const FieldInfo& f) {
    TileInfo ti;
    int bid = tile;
    const int tile_x = bid % a.ntx;
    bid /= a.ntx;
    const int tile_y = bid % a.nty;
    const int tile_z = bid / a.nty;
    ti.x0 = tile_x * 32;
    ti.y0 = tile_y * 8;
    ti.zt = a.zt0 + tile_z * 8;
    ti.near_mask = a.tile_flags[tile] & 15;
    ti.sat_mask = (a.tile_flags[tile] >> 8) & 15;
    ti.translation_only = f.translation_only;
    ti.reff2 = f.reff2;
    ti.r_brick = 3.5f * sqrtf(a.vsx * a.vsx + a.vsy * a.vsy + a.vsz * a.vsz);  // brick half diagonal
    ti.brick0 = a.warped ? (size_t) (ti.x0 / 8) + (size_t) a.bdx * ((ti.y0 / 8) + (size_t) a.bdy * (ti.zt / 8)) : 0;
    return ti;
}

// streaming update of a quad all of whose voxels receive tsdf = 1 (saturated free space): no projection, no depth fetch.
// The store is elided when nothing changes (a voxel that already holds (1.0, max_weight) stays what it is).
DFU_DEV void quad_saturate(const IntegrateArgs& a, size_t lin, Tally& tally) {
    uint4* p = reinterpret_cast<uint4*>(a.vol + lin);
    const uint4 v = ld_stream(p);
    uint4 n;
    n.x = voxel_update_one(a, v.x);
    n.y = voxel_update_one(a, v.y);
    n.z = voxel_update_one(a, v.z);
    n.w = voxel_update_one(a, v.w);
    tally.vox += 4;
    tally.quads += 1;
    tally.sat += 4;
    if ((n.x ^ v.x) | (n.y ^ v.y) | (n.z ^ v.z) | (n.w ^ v.w)) st_stream(p, n);
}

// an entirely saturated tile (every voxel receives tsdf = 1): each thread its four quads with all four 128-bit loads in
// flight; stores only where a voxel changes.  These tiles travel in the same ticket list as the others, so their streaming
// overlaps the latency-bound near bricks of the neighbouring CTAs.
DFU_DEV void saturate_tile(const IntegrateArgs& a, const TileInfo& ti, Tally& tally) {
    const size_t plane = (size_t) a.dx * a.dy;
    uint4* p[4];
    uint4 v[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int lin = it * 128 + threadIdx.x;
        const int qx = lin & 7, yy = (lin >> 3) & 7, zz = lin >> 6;
        p[it] = reinterpret_cast<uint4*>(a.vol + (size_t) (ti.x0 + qx * 4) + (size_t) (ti.y0 + yy) * a.dx + plane * (size_t) (ti.zt + zz));
        v[it] = ld_stream(p[it]);
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        uint4 nv;
        nv.x = voxel_update_one(a, v[it].x);
        nv.y = voxel_update_one(a, v[it].y);
        nv.z = voxel_update_one(a, v[it].z);
        nv.w = voxel_update_one(a, v[it].w);
        if ((nv.x ^ v[it].x) | (nv.y ^ v[it].y) | (nv.z ^ v[it].z) | (nv.w ^ v[it].w)) st_stream(p[it], nv);
    }
    tally.vox += 16;
    tally.quads += 4;
    tally.sat += 16;
}

// rigid pass over the quads of the bricks that are not near (saturated bricks: streaming update)
DFU_DEV void rigid_pass(const IntegrateArgs& a, const TileInfo& ti, bool skip_rigid, Tally& tally) {
    const size_t plane = (size_t) a.dx * a.dy;
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        const int lin = it * 128 + threadIdx.x;
        const int qx = lin & 7, yy = (lin >> 3) & 7, zz = lin >> 6;
        const int z = ti.zt + zz;
        const int sb = qx >> 1;
        if (z < a.z0 || z >= a.z1) continue;
        const int x = ti.x0 + qx * 4, y = ti.y0 + yy;
        const size_t vlin = (size_t) x + (size_t) y * a.dx + plane * (size_t) z;
        if ((ti.sat_mask >> sb) & 1) {
            quad_saturate(a, vlin, tally);
            continue;
        }
        if (skip_rigid || ((ti.near_mask >> sb) & 1)) continue;
        const float py = fmul((float) y, a.vsy), pz = fmul((float) z, a.vsz);
        bool hit[4];
        float ts[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) hit[v] = voxel_tsdf(a, fmul((float) (x + v), a.vsx), py, pz, ts[v]);
        quad_commit(a, vlin, hit, ts, tally);
    }
}

// exact 8-NN of the thread's 4 voxels of brick sb: cull the N nodes to the brick's candidates (every warp compacts
// its slice in index order with ballots), then scan them with the nanoflann metric.  All 128 threads take part.
DFU_DEV void brick_knn_scan(const IntegrateArgs& a, const TileInfo& ti, IntegrateSmem& sm, int sb, const float (&px)[4], float py,
                            float pz, Top8 (&t)[4]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float bcx = (ti.x0 + sb * 8 + 3.5f) * a.vsx, bcy = (ti.y0 + 3.5f) * a.vsy, bcz = (ti.zt + 3.5f) * a.vsz;
    const float d8c = sqrtf(__ldg(&a.bounds[ti.brick0 + sb]).x);
    // any node among the 8 nearest of ANY voxel of the brick lies within d8(centre) + 2r of the centre
    const float th = (d8c + 2.f * ti.r_brick) * 1.0001f + 1e-6f;
    const float thr2 = th * th;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        // the 8 nodes nearest to the brick centre are within d8c + |v - centre| of voxel v: start every
        // slot at that (inflated) bound, so the bulk of the candidates fails the first compare
        const float ex = px[v] - bcx, ey = py - bcy, ez = pz - bcz;
        const float bnd = (d8c + sqrtf(ex * ex + ey * ey + ez * ez)) * 1.0001f + 1e-6f;
        const float bnd2 = bnd * bnd;
#pragma unroll
        for (int k = 0; k < DFU_KNN; ++k) {
            t[v].d[k] = bnd2;
            t[v].i[k] = -1;
        }
    }
#pragma unroll 1
    for (int c0 = 0; c0 < a.N; c0 += CHUNK) {
        int n = 0;
#pragma unroll 1
        for (int j = 0; j < SEG; j += 32) {
            const int idx = c0 + warp * SEG + j + lane;
            bool keep = false;
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < a.N) {
                p = __ldg(&a.pos_w[idx]);
                const float ddx = p.x - bcx, ddy = p.y - bcy, ddz = p.z - bcz;
                keep = (ddx * ddx + ddy * ddy + ddz * ddz) <= thr2;
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                p.w = __int_as_float(idx);
                sm.cand[warp * SEG + n + __popc(m & ((1u << lane) - 1u))] = p;
            }
            n += __popc(m);
        }
        if (lane == 0) sm.cnt[warp] = n;
        __syncthreads();
#pragma unroll 1
        for (int sg = 0; sg < 4; ++sg) {
            const int cn = sm.cnt[sg];
            const float4* __restrict__ cl = sm.cand + sg * SEG;
#pragma unroll 1
            for (int j = 0; j < cn; ++j) {
                const float4 p = cl[j];  // warp-wide broadcast
                const int idx = __float_as_int(p.w);
                const float dy = fsub(py, p.y), dz = fsub(pz, p.z);
                const float dy2 = fmul(dy, dy), dz2 = fmul(dz, dz);
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float dxv = fsub(px[v], p.x);
                    const float d = fadd(fadd(fmul(dxv, dxv), dy2), dz2);  // nanoflann metric, bit exact
                    if (d < t[v].d[DFU_KNN - 1]) top8_insert(t[v], d, idx);
                }
            }
        }
        __syncthreads();
    }
}

// warped position of one voxel from its 8 neighbour ids (ascending by (dist2, idx)); distances are re-evaluated
// with the scan's expression.  Translation-only fields take the reduced chain (blend.cuh), everything else the
// general blend.  The 8 FP64 weights are evaluated as independent straight-line chains.
DFU_DEV V3 warp_voxel(const IntegrateArgs& a, const TileInfo& ti, const int (&id)[DFU_KNN], float px, float py, float pz,
                      bool off_zero_planes) {
    float4 nd[DFU_KNN];
    float d2[DFU_KNN];
#pragma unroll
    for (int k = 0; k < DFU_KNN; ++k) {
        nd[k] = id[k] >= 0 ? __ldg(&a.pos_w[id[k]]) : make_float4(INFINITY, INFINITY, INFINITY, 1.f);
        d2[k] = dist2(px, py, pz, nd[k].x, nd[k].y, nd[k].z);
    }
    const bool fast = ti.translation_only && a.blend_mode == DFU_BLEND_REF_COMPOSE;
    if (fast && d2[0] > ti.reff2 && off_zero_planes) return V3{px, py, pz};  // rule (b): p' == p bit for bit
    float w[DFU_KNN];
    double arg[DFU_KNN];
    bool zero[DFU_KNN];
#pragma unroll
    for (int k = 0; k < DFU_KNN; ++k) {  // Node::getTransformationWeight (node.cpp:29-36), see dfu_math.cuh node_weight
        zero[k] = id[k] < 0 || d2[k] > fmul(209.f, fmul(nd[k].w, nd[k].w));
        const double dx = (double) fsub(nd[k].x, px), dy = (double) fsub(nd[k].y, py), dz = (double) fsub(nd[k].z, pz);
        const double distSq = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        const double w2 = __dmul_rn((double) nd[k].w, (double) nd[k].w);
        arg[k] = zero[k] ? 0.0 : __ddiv_rn(-distSq, __dmul_rn(2.0, w2));
    }
#pragma unroll
    for (int k = 0; k < DFU_KNN; ++k) w[k] = zero[k] ? 0.f : __double2float_rn(exp(arg[k]));
    if (fast) {
        float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
        for (int k = 0; k < DFU_KNN; ++k) {
            if (id[k] < 0) break;
            if (w[k] != 0.f) {
                const float4 du = __ldg(&a.dual[id[k]]);  // (w,x,y,z) stored in (x,y,z,w)
                ax = fadd(fmul(du.y, w[k]), ax);
                ay = fadd(fmul(du.z, w[k]), ay);
                az = fadd(fmul(du.w, w[k]), az);
            }
        }
        return V3{fadd(px, fmul(2.f, ax)), fadd(py, fmul(2.f, ay)), fadd(pz, fmul(2.f, az))};
    }
    Top8 t;
#pragma unroll
    for (int k = 0; k < DFU_KNN; ++k) {
        t.i[k] = id[k];
        t.d[k] = d2[k];
    }
    const DQ b = blend(a.blend_mode, t, w, a.real, a.dual);
    return dq_transform_vertex(b, V3{px, py, pz});
}

// warped position of one voxel from its cached 8 neighbour ids and weights (no distance or weight evaluation:
// both depend on positions only and were computed when the brick was filled)
DFU_DEV V3 warp_voxel_weights(const IntegrateArgs& a, const TileInfo& ti, const int (&id)[DFU_KNN], const float (&w)[DFU_KNN],
                              float px, float py, float pz) {
    if (ti.translation_only && a.blend_mode == DFU_BLEND_REF_COMPOSE) {
        float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
        for (int k = 0; k < DFU_KNN; ++k) {
            if (id[k] >= 0 && w[k] != 0.f) {
                const float4 du = __ldg(&a.dual[id[k]]);  // (w,x,y,z) stored in (x,y,z,w)
                ax = fadd(fmul(du.y, w[k]), ax);
                ay = fadd(fmul(du.z, w[k]), ay);
                az = fadd(fmul(du.w, w[k]), az);
            }
        }
        return V3{fadd(px, fmul(2.f, ax)), fadd(py, fmul(2.f, ay)), fadd(pz, fmul(2.f, az))};
    }
    Top8 t;
#pragma unroll
    for (int k = 0; k < DFU_KNN; ++k) {
        t.i[k] = id[k];
        t.d[k] = 0.f;  // not used by blend() once the weights are known
    }
    const DQ b = blend(a.blend_mode, t, w, a.real, a.dual);
    return dq_transform_vertex(b, V3{px, py, pz});
}

DFU_DEV uint4 pack_ids(const Top8& t) {
    unsigned u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = ((unsigned) t.i[2 * k] & 0xffffu) | (((unsigned) t.i[2 * k + 1] & 0xffffu) << 16);
    return make_uint4(u[0], u[1], u[2], u[3]);
}
DFU_DEV void unpack_ids(uint4 c, int (&id)[DFU_KNN]) {
    const unsigned u[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int k = 0; k < DFU_KNN; ++k) {
        const int v = (int) ((u[k >> 1] >> ((k & 1) * 16)) & 0xffffu);
        id[k] = v == 0xffff ? -1 : v;
    }
}

enum { MODE_ONTHEFLY = 0, MODE_FILL = 1, MODE_CACHED = 2 };

// MODE_ONTHEFLY: rigid pass + per-voxel 8-NN recomputed for every near brick (no cache)
// MODE_FILL    : only computes and stores the 8-NN of near bricks that are not in the cache yet (no TSDF work)
// MODE_CACHED  : rigid pass + near bricks from the cache (filled by a MODE_FILL launch earlier in the stream);
//                no per-thread top-8 lists -> half the registers, twice the resident warps
template <int MODE>
__global__ void __launch_bounds__(128, MODE == MODE_CACHED ? 8 : 4) integrate_kernel(const __grid_constant__ IntegrateArgs a) {
    const int tid = threadIdx.x;
    const FieldInfo fi = field_info(a);
    const size_t plane = (size_t) a.dx * a.dy;
    const int n_work = a.work_count[MODE == MODE_FILL ? 1 : 0];
    const int* __restrict__ list = MODE == MODE_FILL ? a.fill_tiles : a.work_tiles;
    const int qx2 = tid & 1, yy = (tid >> 1) & 7, zz = tid >> 4;
    // persistent CTAs pull tiles from the compacted list of tiles that have work (tile_classify_kernel); a shared
    // ticket counter balances the very uneven tiles (rigid-only vs near bricks)
    __shared__ int s_ticket[2];
    Tally tally;
    // the ticket of the NEXT tile is drawn while the current one is processed (an atomic round trip per tile otherwise)
    if (tid == 0) s_ticket[0] = atomicAdd(&a.work_count[MODE == MODE_FILL ? 3 : 2], 1);
    int par = 0;
#pragma unroll 1
    for (;;) {
        __syncthreads();
        const int wi = s_ticket[par];
        if (wi >= n_work) break;
        int next_tile = -1;
        if (tid == 0) {
            const int nwi = atomicAdd(&a.work_count[MODE == MODE_FILL ? 3 : 2], 1);
            s_ticket[par ^ 1] = nwi;
            if (MODE == MODE_CACHED && nwi < n_work) next_tile = list[nwi];  // (consumed after the rigid pass: the load has time to land)
        }
        par ^= 1;
        const int tile = list[wi];
        const TileInfo ti = tile_info(a, tile, fi);
        if (MODE != MODE_FILL && ti.sat_mask == 15 && ti.zt >= a.z0 && ti.zt + 8 <= a.z1) {
            saturate_tile(a, ti, tally);
            continue;
        }
        if (MODE != MODE_FILL && (!(a.tile_flags[tile] & 16) || ti.sat_mask)) rigid_pass(a, ti, (a.tile_flags[tile] & 16) != 0, tally);
        if (MODE == MODE_CACHED && next_tile >= 0) {
            // the neighbour-cache slabs (8 KB of ids + 16 KB of weights per brick, contiguous) of the NEXT tile's near bricks are
            // pulled into L2 by the TMA engine while this tile's bricks are being processed: the per-voxel cache reads were the
            // top stall of this kernel (dependent DRAM loads at 50 % occupancy)
            const int nm = a.tile_flags[next_tile] & 15;
            if (nm) {
                int b = next_tile;
                const int ntx_ = b % a.ntx;
                b /= a.ntx;
                const int nty_ = b % a.nty, ntz_ = b / a.nty;
                const size_t nb0 = (size_t) (ntx_ * 4) + (size_t) a.bdx * ((size_t) nty_ + (size_t) a.bdy * (size_t) ((a.zt0 >> 3) + ntz_));
#pragma unroll
                for (int sb = 0; sb < 4; ++sb)
                    if ((nm >> sb) & 1) {
                        tma_prefetch_l2(a.knn_pool + (nb0 + sb) * 512, 512 * 16);
                        tma_prefetch_l2(a.w_pool + (nb0 + sb) * 1024, 1024 * 16);
                    }
            }
        }
        if (ti.near_mask == 0) continue;
        const int z = ti.zt + zz, y = ti.y0 + yy;
        const float py = fmul((float) y, a.vsy), pz = fmul((float) z, a.vsz);
#pragma unroll 1
        for (int sb = 0; sb < 4; ++sb) {
            if (!((ti.near_mask >> sb) & 1)) continue;  // uniform over the CTA
            const int x = ti.x0 + sb * 8 + qx2 * 4;
            const size_t brick = ti.brick0 + sb;
            const size_t lin = (size_t) x + (size_t) y * a.dx + plane * (size_t) z;
            float px[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) px[v] = fmul((float) (x + v), a.vsx);

            if (MODE == MODE_CACHED) {
                if (z < a.z0 || z >= a.z1) continue;
                const size_t slot = brick * 512 + ((zz * 8 + yy) * 8 + qx2 * 4);
                const uint4* cache = a.knn_pool + slot;
                const float4* wcache = a.w_pool + 2 * slot;
                bool hit[4];
                float ts[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    int id[DFU_KNN];
                    // (48 bytes per voxel read exactly once: streamed past the L1, which keeps the nodes' dual parts and the
                    //  depth image that neighbouring voxels share)
                    unpack_ids(__ldcs(cache + v), id);
                    float w[DFU_KNN];
                    ld256_cs(reinterpret_cast<const float*>(wcache + 2 * v), w);
                    const V3 p = warp_voxel_weights(a, ti, id, w, px[v], py, pz);
                    hit[v] = voxel_tsdf(a, p.x, p.y, p.z, ts[v]);
                }
                quad_commit(a, lin, hit, ts, tally);
                if (tid == 0) tally.bricks += 1;
            } else {
                __shared__ IntegrateSmem sm;
                if (MODE == MODE_FILL && a.built[brick]) continue;  // uniform over the CTA
                Top8 t[4];
                brick_knn_scan(a, ti, sm, sb, px, py, pz, t);
                if (MODE == MODE_FILL) {
                    // 8 u16 ids per voxel, 64 contiguous bytes per thread
                    const size_t slot = brick * 512 + ((zz * 8 + yy) * 8 + qx2 * 4);
                    uint4* cache = a.knn_pool + slot;
                    float4* wcache = a.w_pool + 2 * slot;
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        cache[v] = pack_ids(t[v]);
                        float w[DFU_KNN];
                        neighbour_weights(t[v], px[v], py, pz, a.pos_w, w);  // the FP64 evaluation, once per voxel
                        wcache[2 * v] = make_float4(w[0], w[1], w[2], w[3]);
                        wcache[2 * v + 1] = make_float4(w[4], w[5], w[6], w[7]);
                    }
                    __syncthreads();
                    if (tid == 0) a.built[brick] = 1;
                } else {
                    if (z < a.z0 || z >= a.z1) continue;
                    bool hit[4];
                    float ts[4];
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const V3 w = warp_voxel(a, ti, t[v].i, px[v], py, pz, x + v > 0 && y > 0 && z > 0);
                        hit[v] = voxel_tsdf(a, w.x, w.y, w.z, ts[v]);
                    }
                    quad_commit(a, lin, hit, ts, tally);
                    if (tid == 0) tally.bricks += 1;
                }
            }
        }
    }
    if (MODE != MODE_FILL && a.stats) {  // one atomic per warp and counter
        unsigned v = tally.vox, q = tally.quads, b = tally.bricks, sa = tally.sat;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
        }
        if ((tid & 31) == 0) {
            if (v) atomicAdd(&a.stats[0], (unsigned long long) v);
            if (q) atomicAdd(&a.stats[1], (unsigned long long) q);
            if (b) atomicAdd(&a.stats[3], (unsigned long long) b);
            if (sa) atomicAdd(&a.stats[2], (unsigned long long) sa);
        }
    }
}

// Classification of the index box [x0,x0+nx-1] x [y0,y0+7] x [z0,z0+7] whose voxels are moved by at most `delta` metres
// (Euclidean) from their grid positions.  Returns bit 0 (BOX_UNREACHABLE) when NO voxel can pass the per-voxel tests of
// voxel_tsdf -- the box (inflated by delta) is behind the camera, projects outside the image, projects only onto pixels
// without depth, or lies entirely more than trunc behind the farthest depth it can see -- and bit 1 (BOX_SATURATED) when EVERY
// voxel passes them with tsdf == 1.0f exactly: the box is in front of the camera, projects inside the image onto pixels
// that all carry a depth, and the nearest of those depths lies more than the truncation distance behind the farthest point of
// the box (sdf >= trunc => fminf(1, sdf / trunc) == 1).  Free space between the camera and the surfaces -- most updated
// voxels of a dense depth image -- then needs no projection at all.  All bounds carry margins far above the rounding of the
// per-voxel arithmetic.
enum { BOX_UNREACHABLE = 1, BOX_SATURATED = 2 };
DFU_DEV int box_test(const IntegrateArgs& a, int x0, int nx, int y0, int z0, float delta) {
    float zmin = INFINITY, zmax = -INFINITY, umin = INFINITY, umax = -INFINITY, vmin = INFINITY, vmax = -INFINITY;
    float cxs = 0.f, cys = 0.f, czs = 0.f, cor[8][3];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float px = (float) (x0 + ((c & 1) ? nx - 1 : 0)) * a.vsx, py = (float) (y0 + ((c & 2) ? 7 : 0)) * a.vsy,
                    pz = (float) (z0 + ((c & 4) ? 7 : 0)) * a.vsz;
        cor[c][0] = a.R[0] * px + a.R[1] * py + a.R[2] * pz + a.T[0];
        cor[c][1] = a.R[3] * px + a.R[4] * py + a.R[5] * pz + a.T[1];
        cor[c][2] = a.R[6] * px + a.R[7] * py + a.R[8] * pz + a.T[2];
        cxs += cor[c][0]; cys += cor[c][1]; czs += cor[c][2];
        zmin = fminf(zmin, cor[c][2]);
        zmax = fmaxf(zmax, cor[c][2]);
    }
    if (zmax + delta <= -1e-4f) return BOX_UNREACHABLE;  // every voxel has vc.z <= 0
    if (!(zmin - delta > 1e-3f)) return 0;
    cxs *= 0.125f; cys *= 0.125f; czs *= 0.125f;
    float rad = 0.f, slope = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float iz = 1.f / cor[c][2];
        const float u = a.fx * cor[c][0] * iz + a.cx, v = a.fy * cor[c][1] * iz + a.cy;
        umin = fminf(umin, u); umax = fmaxf(umax, u);
        vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
        slope = fmaxf(slope, fmaxf(fabsf(cor[c][0]), fabsf(cor[c][1])) * iz);
        const float ex = cor[c][0] - cxs, ey = cor[c][1] - cys, ez = cor[c][2] - czs;
        rad = fmaxf(rad, sqrtf(ex * ex + ey * ey + ez * ez));
    }
    // a displacement e, |e| <= delta, moves the projection by at most f * delta * (1 + |x|/z) / (z - delta) pixels
    const float mpx = delta > 0.f ? fmaxf(a.fx, a.fy) * delta * (1.f + slope) / (zmin - delta) * 1.01f : 0.f;
    // pixel rectangle the box can project to, one pixel of margin
    const int pu0 = max(0, (int) floorf(fmaxf(umin - mpx, -1e6f)) - 1), pu1 = min(a.cols - 1, (int) floorf(fminf(umax + mpx, 1e6f)) + 1);
    const int pv0 = max(0, (int) floorf(fmaxf(vmin - mpx, -1e6f)) - 1), pv1 = min(a.rows - 1, (int) floorf(fminf(vmax + mpx, 1e6f)) + 1);
    if (pu0 > pu1 || pv0 > pv1) return BOX_UNREACHABLE;  // projects outside the image
    float dfar = 0.f, dnear = INFINITY;
    for (int ty = pv0 / DT; ty <= pv1 / DT; ++ty)
        for (int tx = pu0 / DT; tx <= pu1 / DT; ++tx) {
            const float2 t = __ldg(&a.dmax_tiles[ty * a.dtx + tx]);
            dfar = fmaxf(dfar, t.x);
            dnear = fminf(dnear, t.y);
        }
    const float cdist = sqrtf(cxs * cxs + cys * cys + czs * czs);
    // dfar == 0: no depth anywhere it projects to; else sdf = Dp - |vc| <= dfar - (|centre| - rad - delta) < -trunc
    if (dfar == 0.f || dfar - (cdist - rad - delta) < -a.trunc - 1e-3f) return BOX_UNREACHABLE;
    // saturated: inside the image by 0.01 pixel (far above the rounding of the per-voxel projection), every pixel valid,
    // sdf >= dnear - (|centre| + rad + delta) >= trunc (1 + 1e-3)
    const bool inside = umin - mpx >= 0.01f && vmin - mpx >= 0.01f && umax + mpx < (float) a.cols - 0.01f && vmax + mpx < (float) a.rows - 0.01f;
    if (inside && dnear > 0.f && dnear < INFINITY && dnear - (cdist + rad + delta) >= a.trunc * 1.001f + 1e-4f) return BOX_SATURATED;
    return 0;
}

// Hierarchical cull of the RIGID part of every tile, one thread per tile.  The tile's un-warped voxels lie in the
// convex hull of its 8 transformed corners.  If that hull is behind the camera, projects outside the image, projects
// only onto pixels without depth, or lies entirely more than trunc behind the farthest depth it can see, no voxel of
// it can pass the per-voxel tests (tsdf_volume.cu:70-79) and the rigid pass is skipped.  All bounds carry margins far
// above the rounding of the per-voxel arithmetic, so the result is bit-identical.
// The same thread classifies the tile's 4 bricks as near / not near (rules (a),(b) above) or saturated (box_test), and
// appends the tile to the work list (anything to do) and to the fill list (a near brick missing from the 8-NN cache).
__global__ void tile_classify_kernel(const __grid_constant__ IntegrateArgs a) {
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= a.ntiles) return;
    if (a.stats && tile < 4) a.stats[tile] = 0ull;  // (the integrate kernels of this call run after this kernel)
    int bid = tile;
    const int x0 = (bid % a.ntx) * 32;
    bid /= a.ntx;
    const int y0 = (bid % a.nty) * 8;
    const int zt = a.zt0 + (bid / a.nty) * 8;
    const int tt = box_test(a, x0, 32, y0, zt, 0.f);
    const bool rigid_skip = (tt & BOX_UNREACHABLE) != 0;
    int near_mask = 0, need_fill = 0, sat_mask = 0;
    if (!a.warped) {
        if (tt & BOX_SATURATED) {
            sat_mask = 15;
        } else if (!rigid_skip) {
#pragma unroll
            for (int sb = 0; sb < 4; ++sb)
                if (box_test(a, x0 + sb * 8, 8, y0, zt, 0.f) & BOX_SATURATED) sat_mask |= 1 << sb;
        }
    }
    if (a.warped) {
        const FieldInfo f = field_info(a);
        const float r_brick = 3.5f * sqrtf(a.vsx * a.vsx + a.vsy * a.vsy + a.vsz * a.vsz);
        const size_t brick0 = (size_t) (x0 / 8) + (size_t) a.bdx * ((y0 / 8) + (size_t) a.bdy * (zt / 8));
#pragma unroll
        for (int sb = 0; sb < 4; ++sb) {
            const float2 b = __ldg(&a.bounds[brick0 + sb]);
            const float dmin = sqrtf(b.y) - r_brick;  // lower bound of voxel-to-node distance in this brick
            const bool on_zero_plane = (x0 + sb * 8 == 0) || (y0 == 0) || (zt == 0);
            if (f.all_near || dmin <= (on_zero_plane ? f.r_zero : f.r_eff)) {
                // translation-only field: a voxel moves by |2 acc_c| <= 16 dmax w per coordinate, w <= exp(-dmin^2 / (2 maxw^2));
                // a brick none of whose moved voxels can be updated needs neither the cache nor the warp, and neither does
                // one all of whose moved voxels are saturated free space
                if (f.translation_only && a.blend_mode == DFU_BLEND_REF_COMPOSE) {
                    const float dm = fmaxf(dmin, 0.f);
                    const float wmax = fminf(1.f, expf(-dm * dm / (2.f * f.maxw * f.maxw)) * 1.001f);
                    const float delta = 1.7321f * 16.f * f.dmax * wmax * 1.001f + 1e-6f;
                    const int bt = box_test(a, x0 + sb * 8, 8, y0, zt, delta);
                    if (bt & BOX_UNREACHABLE) continue;
                    if (bt & BOX_SATURATED) {
                        sat_mask |= 1 << sb;
                        continue;
                    }
                }
                near_mask |= 1 << sb;
                if (a.knn_pool && !a.built[brick0 + sb]) need_fill = 1;
            } else if (tt & BOX_SATURATED) {
                sat_mask |= 1 << sb;  // un-warped brick of a saturated tile
            } else if (!rigid_skip && (box_test(a, x0 + sb * 8, 8, y0, zt, 0.f) & BOX_SATURATED)) {
                sat_mask |= 1 << sb;  // un-warped brick in saturated free space
            }
        }
    }
    a.tile_flags[tile] = (unsigned short) (near_mask | (rigid_skip ? 16 : 0) | (sat_mask << 8));
    if (near_mask || sat_mask || !rigid_skip) a.work_tiles[atomicAdd(&a.work_count[0], 1)] = tile;  // (work_count[0] is copied to stats[2] by the host entry)
    if (need_fill) a.fill_tiles[atomicAdd(&a.work_count[1], 1)] = tile;
}

// max and min ray length per 16x16-pixel tile of the dists image (for the hierarchical cull / the saturation test)
__global__ void __launch_bounds__(DT * DT) depth_tiles_kernel(const uint16_t* __restrict__ dists, size_t pitch, int rows, int cols,
                                                              float2* __restrict__ out, int dtx) {
    __shared__ float sh[2][DT * DT / 32];
    const int x = blockIdx.x * DT + (threadIdx.x % DT), y = blockIdx.y * DT + (threadIdx.x / DT);
    float d = 0.f, dn = INFINITY;  // pixels beyond the image border do not lower the minimum
    if (x < cols && y < rows) {
        d = __half2float(__ushort_as_half(*reinterpret_cast<const unsigned short*>(reinterpret_cast<const char*>(dists) + (size_t) y * pitch + 2 * (size_t) x)));
        dn = d;
    }
    if (!(d == d)) {  // NaN depth: never cull, never saturate
        d = INFINITY;
        dn = 0.f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        d = fmaxf(d, __shfl_xor_sync(0xffffffffu, d, o));
        dn = fminf(dn, __shfl_xor_sync(0xffffffffu, dn, o));
    }
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = d;
        sh[1][threadIdx.x >> 5] = dn;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.f, mn = INFINITY;
        for (int i = 0; i < DT * DT / 32; ++i) {
            m = fmaxf(m, sh[0][i]);
            mn = fminf(mn, sh[1][i]);
        }
        out[blockIdx.y * dtx + blockIdx.x] = make_float2(m, mn);
    }
}

// compute_dists_kernel (src/kfusion/cuda/imgproc.cu:233-245); the reference's guard uses || (:237, a
// latent out-of-bounds access) -- the bounds check here is the intended &&.
__global__ void compute_dists_kernel(const uint16_t* __restrict__ depth, size_t dpitch, uint16_t* __restrict__ dists,
                                     size_t opitch, int rows, int cols, float finvx, float finvy, float cx, float cy) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < cols && y < rows) {
        const float xl = fmul(fsub((float) x, cx), finvx);
        const float yl = fmul(fsub((float) y, cy), finvy);
        const float lambda = __fsqrt_rn(fadd(fadd(fmul(xl, xl), fmul(yl, yl)), 1.f));
        const uint16_t d = *reinterpret_cast<const uint16_t*>(reinterpret_cast<const char*>(depth) + (size_t) y * dpitch + 2 * (size_t) x);
        *reinterpret_cast<uint16_t*>(reinterpret_cast<char*>(dists) + (size_t) y * opitch + 2 * (size_t) x) =
            __half_as_ushort(__float2half_rn(fmul(fmul((float) d, lambda), 0.001f)));
    }
}

}  // namespace

// Private stream-ordered pool for the per-call scratch (depth tiles + cull flags).  The default pool hands its
// memory back at every synchronisation (release threshold 0), which would turn each call after a sync into a real
// allocation; this one keeps up to 64 MiB cached and leaves the application's pools alone.
cudaMemPool_t scratch_pool(int device) {
    static cudaMemPool_t pools[64] = {};
    if (device < 0 || device >= 64) return nullptr;
    if (!pools[device]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t p = nullptr;
        if (cudaMemPoolCreate(&p, &props) != cudaSuccess) {
            (void) cudaGetLastError();
            return nullptr;
        }
        unsigned long long keep = 64ull << 20;
        cudaMemPoolSetAttribute(p, cudaMemPoolAttrReleaseThreshold, &keep);
        pools[device] = p;
    }
    return pools[device];
}

// instrumentation of the last dfu_tsdf_integrate call per device (see IntegrateArgs::stats)
static unsigned long long* integrate_stats(int device) {
    static unsigned long long* bufs[64] = {};
    if (device < 0 || device >= 64) return nullptr;
    if (!bufs[device]) {
        if (cudaMalloc(&bufs[device], 8 * sizeof(unsigned long long)) != cudaSuccess) {
            (void) cudaGetLastError();
            bufs[device] = nullptr;
        } else {
            cudaMemset(bufs[device], 0, 8 * sizeof(unsigned long long));
        }
    }
    return bufs[device];
}

extern "C" {

int dfu_tsdf_integrate_stats(unsigned long long stats_host[4], dfu_stream stream) {
    DFU_REQUIRE(stats_host, DFU_ERR_INVALID, "NULL argument");
    int device = 0;
    DFU_CUDA_OK(cudaGetDevice(&device));
    unsigned long long* d = integrate_stats(device);
    DFU_REQUIRE(d != nullptr, DFU_ERR_CUDA, "no statistics buffer");
    DFU_CUDA_OK(cudaMemcpyAsync(stats_host, d, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, as_stream(stream)));
    DFU_CUDA_OK(cudaStreamSynchronize(as_stream(stream)));
    return DFU_OK;
}

int dfu_compute_dists(const uint16_t* depth, size_t depth_pitch_bytes, uint16_t* dists, size_t dists_pitch_bytes,
                      int rows, int cols, const float intr_host[4], dfu_stream stream) {
    DFU_REQUIRE(depth && dists && intr_host, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(rows > 0 && cols > 0, DFU_ERR_INVALID, "bad image size");
    DFU_GUARD(dfu_device_of(depth));
    dim3 block(32, 8), grid(div_up(cols, 32), div_up(rows, 8));
    compute_dists_kernel<<<grid, block, 0, as_stream(stream)>>>(depth, depth_pitch_bytes, dists, dists_pitch_bytes, rows,
                                                                cols, 1.f / intr_host[0], 1.f / intr_host[1],
                                                                intr_host[2], intr_host[3]);
    DFU_LAUNCH_OK();
    return DFU_OK;
}

float dfu_tsdf_trunc_dist(float requested, const float vs[3]) {
    const float m = fmaxf(fmaxf(vs[0], vs[1]), vs[2]);
    return fmaxf(requested, 2.1f * m);
}

int dfu_tsdf_clear(void* volume, const int dims[3], int z0, int z1, dfu_stream stream) {
    DFU_REQUIRE(volume && dims, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(0 <= z0 && z0 <= z1 && z1 <= dims[2], DFU_ERR_INVALID, "bad z range");
    DFU_GUARD(dfu_device_of(volume));
    const size_t plane = (size_t) dims[0] * dims[1];
    // pack_tsdf(0.f, 0) == 0x00000000 (tsdf_volume.cu:20)
    DFU_CUDA_OK(cudaMemsetAsync(reinterpret_cast<uint32_t*>(volume) + plane * (size_t) z0, 0,
                                plane * (size_t) (z1 - z0) * sizeof(uint32_t), as_stream(stream)));
    return DFU_OK;
}

int dfu_tsdf_integrate(void* volume, const int dims[3], const float vs[3], float trunc_dist, int max_weight,
                       const float vol2cam[12], const float intr[4], const uint16_t* dists, size_t pitch, int rows,
                       int cols, dfu_warpfield* wf, int blend_mode, int z0, int z1, dfu_stream stream) {
    DFU_REQUIRE(volume && dims && vs && vol2cam && intr && dists, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(dims[0] > 0 && dims[0] % 32 == 0, DFU_ERR_INVALID, "dims.x must be a positive multiple of 32");
    DFU_REQUIRE(dims[1] > 0 && dims[1] % 8 == 0, DFU_ERR_INVALID, "dims.y must be a positive multiple of 8");
    DFU_REQUIRE(0 <= z0 && z0 <= z1 && z1 <= dims[2], DFU_ERR_INVALID, "bad z range");
    DFU_REQUIRE(max_weight >= 1 && max_weight <= 65535, DFU_ERR_INVALID, "max_weight out of the u16 range");
    DFU_REQUIRE(((uintptr_t) volume & 15) == 0, DFU_ERR_INVALID, "volume must be 16-byte aligned");
    DFU_REQUIRE(blend_mode == DFU_BLEND_REF_COMPOSE || blend_mode == DFU_BLEND_DQB_SUM, DFU_ERR_INVALID, "bad blend_mode");
    if (z0 == z1) return DFU_OK;
    DFU_GUARD(wf ? wf->device : dfu_device_of(volume));  // (also drops stale errors of other libraries)
    cudaStream_t st = as_stream(stream);
    IntegrateArgs a{};
    a.vol = reinterpret_cast<uint32_t*>(volume);
    a.dx = dims[0]; a.dy = dims[1]; a.dz = dims[2];
    a.vsx = vs[0]; a.vsy = vs[1]; a.vsz = vs[2];
    a.trunc = trunc_dist;
    a.trunc_inv = 1.f / trunc_dist;  // tsdf_volume.cu:106
    a.max_weight = max_weight;
    for (int i = 0; i < 9; ++i) a.R[i] = vol2cam[i];
    for (int i = 0; i < 3; ++i) a.T[i] = vol2cam[9 + i];
    a.fx = intr[0]; a.fy = intr[1]; a.cx = intr[2]; a.cy = intr[3];
    a.dists = dists;
    a.pitch = pitch;
    a.rows = rows;
    a.cols = cols;
    a.z0 = z0;
    a.z1 = z1;
    a.zt0 = z0 / 8 * 8;
    a.ntx = dims[0] / 32;
    a.nty = dims[1] / 8;
    const int ntz = (z1 - a.zt0 + 7) / 8;
    if (wf) {
        DFU_REQUIRE(wf->initialised, DFU_ERR_NOT_INIT, "warp field not initialised");
        int rc = dfu_wf_build_brick_table(wf, dims, vs, z0, z1, st);
        if (rc != DFU_OK) return rc;
        a.warped = 1;
        a.blend_mode = blend_mode;
        a.pos_w = wf->pos_w;
        a.real = wf->real;
        a.dual = wf->dual;
        a.N = wf->N;
        a.flags = wf->flags;
        a.bounds = wf->bricks.bounds;
        a.bdx = dims[0] / 8;
        a.bdy = dims[1] / 8;
        if (wf->bricks.knn_pool) {  // pools start at brick plane pool_zb0: shift so that they index by global brick id
            const size_t off = (size_t) (dims[0] / 8) * (dims[1] / 8) * (size_t) wf->bricks.pool_zb0;
            a.knn_pool = wf->bricks.knn_pool - off * 512;
            a.w_pool = wf->bricks.w_pool - off * 1024;
            a.built = wf->bricks.built - off;
        }
    }
    const long nblocks = (long) a.ntx * a.nty * ntz;
    DFU_REQUIRE(nblocks <= 0x7fffffffL, DFU_ERR_INVALID, "volume too large for one launch");
    // per-call scratch (stream-ordered: safe with concurrent streams): coarse max-depth map, per-tile flags, work lists
    a.dtx = div_up(cols, DT);
    a.dty = div_up(rows, DT);
    a.ntiles = (int) nblocks;
    auto up = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t o_flags = up((size_t) a.dtx * a.dty * sizeof(float2));
    const size_t o_count = o_flags + up((size_t) nblocks * sizeof(unsigned short));
    const size_t o_work = o_count + 256;
    const size_t o_fill = o_work + up((size_t) nblocks * sizeof(int));
    const size_t total = o_fill + up((size_t) nblocks * sizeof(int));
    char* tiles = nullptr;
    int device = 0;
    DFU_CUDA_OK(cudaGetDevice(&device));
    cudaMemPool_t pool = scratch_pool(device);
    if (pool)
        DFU_CUDA_OK(cudaMallocFromPoolAsync((void**) &tiles, total, pool, st));
    else
        DFU_CUDA_OK(cudaMallocAsync((void**) &tiles, total, st));
    a.dmax_tiles = reinterpret_cast<float2*>(tiles);
    a.tile_flags = reinterpret_cast<unsigned short*>(tiles + o_flags);
    a.work_count = reinterpret_cast<int*>(tiles + o_count);
    a.work_tiles = reinterpret_cast<int*>(tiles + o_work);
    a.fill_tiles = reinterpret_cast<int*>(tiles + o_fill);
    DFU_CUDA_OK(cudaMemsetAsync(a.work_count, 0, 8 * sizeof(int), st));
    a.stats = integrate_stats(device);
    depth_tiles_kernel<<<dim3(a.dtx, a.dty), DT * DT, 0, st>>>(dists, pitch, rows, cols, reinterpret_cast<float2*>(tiles), a.dtx);
    DFU_LAUNCH_OK();
    tile_classify_kernel<<<div_up(nblocks, 128), 128, 0, st>>>(a);
    DFU_LAUNCH_OK();
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const unsigned grid8 = (unsigned) std::min<long>(nblocks, (long) sms * 8), grid4 = (unsigned) std::min<long>(nblocks, (long) sms * 4);
    if (!a.warped || a.knn_pool == nullptr) {
        integrate_kernel<MODE_ONTHEFLY><<<grid4, 128, 0, st>>>(a);
        DFU_LAUNCH_OK();
    } else {
        integrate_kernel<MODE_FILL><<<grid4, 128, 0, st>>>(a);  // its work list is empty once the cache is warm
        DFU_LAUNCH_OK();
        integrate_kernel<MODE_CACHED><<<grid8, 128, 0, st>>>(a);
        DFU_LAUNCH_OK();
    }
    DFU_CUDA_OK(cudaFreeAsync(tiles, st));
    return DFU_OK;
}

}  // extern "C"
