// Warpfield::update (src/dynfu/warp_field.cpp:34-95) on the device:
//   getUnsupportedVertices (:34-62)  8-NN of every vertex, min_k dist/dg_w >= 1  -> ordered compaction
//   pcl::VoxelGrid, 5 cm leaf (:68-72) bounding box -> cell keys -> counting sort -> per-cell centroid (ascending
//                                     point index inside a cell, cells in ascending linear index)
//   node insertion (:77-84)           dg_v = centroid, dg_se3 = calcDQB(centroid) against the old nodes, dg_w = 2 eps
//   KD-tree rebuild (:86-94)          re-pack the node arrays, rebuild the node grid; the per-voxel 8-NN cache of the
//                                     integrator is invalidated only for bricks a new node can reach.
// The reference does this with a serial CPU loop over all vertices, PCL and a full nanoflann rebuild.
#include <math_constants.h>

#include <algorithm>

#include "dfu_internal.h"
#include "dfu_math.cuh"
#include "scan.cuh"

using namespace dfu;

namespace {

constexpr int VG_MAX_CELLS = 1 << 21;  // dense cell table (a 6.4 m cube at the reference's 5 cm leaf)
constexpr int VG_CHUNK = 1024;
constexpr int VG_MAX_CHUNKS = VG_MAX_CELLS / VG_CHUNK;

// device-side state of one update / filter call
struct VgState {
    int U;          // input points of the voxel grid (= unsupported vertices)
    int M;          // output points (= non-empty cells)
    int overflow;   // the cell table would not fit
    int ncell;
    unsigned bbox[6];
    unsigned done;
    int min_b[3], div_b[3];
};

// ---- getUnsupportedVertices -------------------------------------------------------------------------------------
__global__ void unsupported_kernel(const float* __restrict__ verts, int P, const int32_t* __restrict__ idx8,
                                   const float4* __restrict__ pos_w, unsigned char* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float vx = verts[3 * (size_t) i], vy = verts[3 * (size_t) i + 1], vz = verts[3 * (size_t) i + 2];
    float mn = CUDART_INF_F;
#pragma unroll
    for (int k = 0; k < DFU_KNN; ++k) {
        const int j = idx8[(size_t) i * DFU_KNN + k];
        if (j < 0) continue;
        const float4 n = __ldg(&pos_w[j]);
        const float dx = fsub(vx, n.x), dy = fsub(vy, n.y), dz = fsub(vz, n.z);
        // sqrt(pow(dx,2)+pow(dy,2)+pow(dz,2)) in double (warp_field.cpp:45-46), stored in a float
        const double s = __dadd_rn(__dadd_rn(__dmul_rn((double) dx, (double) dx), __dmul_rn((double) dy, (double) dy)),
                                   __dmul_rn((double) dz, (double) dz));
        const float dist = __double2float_rn(__dsqrt_rn(s));
        const float r = __fdiv_rn(dist, n.w);
        if (r <= mn) mn = r;
    }
    flags[i] = mn >= 1.f ? 1 : 0;
}

// ordered compaction of the flagged vertices: one warp per 32 vertices
__global__ void flag_count_kernel(const unsigned char* __restrict__ flags, int P, int* __restrict__ counts) {
    const int chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (chunk * 32 >= P) return;
    const int i = chunk * 32 + lane;
    const unsigned m = __ballot_sync(0xffffffffu, i < P && flags[i]);
    if (lane == 0) counts[chunk] = __popc(m);
}
__global__ void flag_emit_kernel(const unsigned char* __restrict__ flags, const float* __restrict__ verts, int P,
                                 const int* __restrict__ offsets, float* __restrict__ out) {
    const int chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (chunk * 32 >= P) return;
    const int i = chunk * 32 + lane;
    const bool ok = i < P && flags[i];
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (!ok) return;
    const size_t slot = (size_t) offsets[chunk] + __popc(m & ((1u << lane) - 1u));
    out[3 * slot] = verts[3 * (size_t) i];
    out[3 * slot + 1] = verts[3 * (size_t) i + 1];
    out[3 * slot + 2] = verts[3 * (size_t) i + 2];
}

// ---- voxel grid ------------------------------------------------------------------------------------------------
__global__ void vg_begin_kernel(VgState* __restrict__ s, int U_or_neg) {
    if (U_or_neg >= 0) s->U = U_or_neg;  // otherwise U was written by the compaction's scan
    s->M = 0;
    s->overflow = 0;
    s->ncell = 0;
    s->done = 0;
    for (int c = 0; c < 3; ++c) {
        s->bbox[c] = 0xffffffffu;
        s->bbox[3 + c] = 0u;
    }
}
__global__ void __launch_bounds__(256) vg_bbox_kernel(const float* __restrict__ pts, VgState* __restrict__ s, float inv_leaf) {
    const int U = s->U;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < U; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = pts[3 * (size_t) i + c];
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            atomicMin(&s->bbox[c], f2ord(mn[c]));
            atomicMax(&s->bbox[3 + c], f2ord(mx[c]));
        }
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&s->done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last || threadIdx.x != 0) return;
    __threadfence();
    if (U <= 0) return;
    long cells = 1;
    for (int c = 0; c < 3; ++c) {  // min_b_ / max_b_ / div_b_ of voxel_grid.hpp
        const float lo = ord2f(atomicOr(&s->bbox[c], 0u)), hi = ord2f(atomicOr(&s->bbox[3 + c], 0u));
        const int mb = (int) floorf(fmul(lo, inv_leaf));
        const int db = (int) floorf(fmul(hi, inv_leaf)) - mb + 1;
        s->min_b[c] = mb;
        s->div_b[c] = db;
        cells = (cells > VG_MAX_CELLS || db > VG_MAX_CELLS) ? (long) VG_MAX_CELLS + 1 : cells * db;
    }
    if (cells > VG_MAX_CELLS) {
        s->overflow = 1;
        s->ncell = 0;
        s->U = 0;  // the remaining kernels become no-ops
    } else {
        s->ncell = (int) cells;
    }
}
__global__ void vg_zero_kernel(const VgState* __restrict__ s, int* __restrict__ count) {
    const int n = s->ncell;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) count[i] = 0;
}
__global__ void vg_count_kernel(const float* __restrict__ pts, const VgState* __restrict__ s, float inv_leaf, int* __restrict__ key,
                                int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s->U) return;
    // ijk = (int)(floor(x * inverse_leaf_size) - (float) min_b)   (voxel_grid.hpp:331-333)
    const int i0 = (int) fsub(floorf(fmul(pts[3 * (size_t) i], inv_leaf)), (float) s->min_b[0]);
    const int i1 = (int) fsub(floorf(fmul(pts[3 * (size_t) i + 1], inv_leaf)), (float) s->min_b[1]);
    const int i2 = (int) fsub(floorf(fmul(pts[3 * (size_t) i + 2], inv_leaf)), (float) s->min_b[2]);
    const int k = i0 + s->div_b[0] * (i1 + s->div_b[1] * i2);
    key[i] = k;
    atomicAdd(&count[k], 1);
}
// exclusive scan over the cells of (points, non-empty flag): per-chunk sums, scan of the sums, per-cell offsets
__global__ void __launch_bounds__(256) vg_chunk_sum_kernel(const VgState* __restrict__ s, const int* __restrict__ count,
                                                           int2* __restrict__ chunk_sum) {
    const int n = s->ncell;
    const int base = blockIdx.x * VG_CHUNK;
    if (base >= n) return;
    int a = 0, b = 0;
#pragma unroll
    for (int k = 0; k < VG_CHUNK / 256; ++k) {
        const int i = base + k * 256 + threadIdx.x;
        const int c = i < n ? count[i] : 0;
        a += c;
        b += c > 0;
    }
    __shared__ int sa[8], sb[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sa[threadIdx.x >> 5] = a;
        sb[threadIdx.x >> 5] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int ta = 0, tb = 0;
        for (int w = 0; w < 8; ++w) {
            ta += sa[w];
            tb += sb[w];
        }
        chunk_sum[blockIdx.x] = make_int2(ta, tb);
    }
}
__global__ void __launch_bounds__(1024) vg_chunk_scan_kernel(VgState* __restrict__ s, int2* __restrict__ chunk_sum) {
    __shared__ int sa[1024], sb[1024];
    const int n = (s->ncell + VG_CHUNK - 1) / VG_CHUNK;
    const int per = (n + 1023) / 1024;
    const int lo = min(n, (int) threadIdx.x * per), hi = min(n, lo + per);
    int a = 0, b = 0;
    for (int i = lo; i < hi; ++i) {
        a += chunk_sum[i].x;
        b += chunk_sum[i].y;
    }
    sa[threadIdx.x] = a;
    sb[threadIdx.x] = b;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int va = (int) threadIdx.x >= o ? sa[threadIdx.x - o] : 0;
        const int vb = (int) threadIdx.x >= o ? sb[threadIdx.x - o] : 0;
        __syncthreads();
        sa[threadIdx.x] += va;
        sb[threadIdx.x] += vb;
        __syncthreads();
    }
    int ra = sa[threadIdx.x] - a, rb = sb[threadIdx.x] - b;
    for (int i = lo; i < hi; ++i) {
        const int2 c = chunk_sum[i];
        chunk_sum[i] = make_int2(ra, rb);
        ra += c.x;
        rb += c.y;
    }
    if (threadIdx.x == 1023) s->M = sb[1023];
}
__global__ void __launch_bounds__(256) vg_offsets_kernel(const VgState* __restrict__ s, const int* __restrict__ count,
                                                         const int2* __restrict__ chunk_sum, int* __restrict__ start,
                                                         int* __restrict__ occ_list) {
    const int n = s->ncell;
    const int base = blockIdx.x * VG_CHUNK;
    if (base >= n) return;
    const int i0 = base + threadIdx.x * 4;  // each thread owns 4 consecutive cells
    int c[4], a = 0, b = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c[k] = i0 + k < n ? count[i0 + k] : 0;
        a += c[k];
        b += c[k] > 0;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int va = __shfl_up_sync(0xffffffffu, ia, o), vb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) {
            ia += va;
            ib += vb;
        }
    }
    __shared__ int wa[8], wb[8];
    if (lane == 31) {
        wa[wid] = ia;
        wb[wid] = ib;
    }
    __syncthreads();
    const int2 cs = chunk_sum[blockIdx.x];
    int ra = cs.x + ia - a, rb = cs.y + ib - b;
    for (int w = 0; w < wid; ++w) {
        ra += wa[w];
        rb += wb[w];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (i0 + k < n) {
            start[i0 + k] = ra;
            if (c[k] > 0) occ_list[rb] = i0 + k;  // non-empty cells in ascending linear index = PCL's output order
        }
        ra += c[k];
        rb += c[k] > 0;
    }
    if (i0 <= n - 1 && n - 1 < i0 + 4) start[n] = ra;
}
__global__ void vg_fill_kernel(const VgState* __restrict__ s, const int* __restrict__ key, const int* __restrict__ start,
                               int* __restrict__ count, int* __restrict__ bucket) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s->U) return;
    const int k = key[i];
    bucket[start[k] + atomicSub(&count[k], 1) - 1] = i;
}
// one warp per non-empty cell: order the cell's points by index (rank sort), add them in that order in float
// (pcl::CentroidPoint / AccumulatorXYZ), divide by the count
__global__ void __launch_bounds__(256) vg_centroid_kernel(const float* __restrict__ pts, const VgState* __restrict__ s,
                                                          const int* __restrict__ start, const int* __restrict__ occ_list,
                                                          const int* __restrict__ bucket, int* __restrict__ bucket2,
                                                          float* __restrict__ out) {
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= s->M) return;
    const int c = occ_list[w];
    const int lo = start[c], n = start[c + 1] - lo;
    for (int i = lane; i < n; i += 32) {
        const int my = bucket[lo + i];
        int r = 0;
        for (int j = 0; j < n; ++j) r += bucket[lo + j] < my;
        bucket2[lo + r] = my;
    }
    __syncwarp();
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (i < n) {
            const size_t id = (size_t) bucket2[lo + i];
            px = pts[3 * id];
            py = pts[3 * id + 1];
            pz = pts[3 * id + 2];
        }
        const int m = min(32, n - base);
        for (int k = 0; k < m; ++k) {
            sx = fadd(sx, __shfl_sync(0xffffffffu, px, k));
            sy = fadd(sy, __shfl_sync(0xffffffffu, py, k));
            sz = fadd(sz, __shfl_sync(0xffffffffu, pz, k));
        }
    }
    if (lane == 0) {
        const float fn = (float) n;
        out[3 * (size_t) w] = __fdiv_rn(sx, fn);
        out[3 * (size_t) w + 1] = __fdiv_rn(sy, fn);
        out[3 * (size_t) w + 2] = __fdiv_rn(sz, fn);
    }
}

struct VgScratch {
    char* base = nullptr;
    VgState* state;
    int *key, *bucket, *bucket2, *occ_list, *count, *start;
    int2* chunk_sum;
};
size_t up256(size_t b) { return (b + 255) / 256 * 256; }
size_t vg_scratch_bytes(int Umax) {
    return up256(sizeof(VgState)) + 4 * up256((size_t) Umax * sizeof(int)) + up256((size_t) VG_MAX_CELLS * sizeof(int)) +
           up256(((size_t) VG_MAX_CELLS + 1) * sizeof(int)) + up256((size_t) VG_MAX_CHUNKS * sizeof(int2));
}
void vg_carve(VgScratch& v, char* p, int Umax) {
    v.base = p;
    v.state = reinterpret_cast<VgState*>(p); p += up256(sizeof(VgState));
    v.key = reinterpret_cast<int*>(p); p += up256((size_t) Umax * sizeof(int));
    v.bucket = reinterpret_cast<int*>(p); p += up256((size_t) Umax * sizeof(int));
    v.bucket2 = reinterpret_cast<int*>(p); p += up256((size_t) Umax * sizeof(int));
    v.occ_list = reinterpret_cast<int*>(p); p += up256((size_t) Umax * sizeof(int));
    v.count = reinterpret_cast<int*>(p); p += up256((size_t) VG_MAX_CELLS * sizeof(int));
    v.start = reinterpret_cast<int*>(p); p += up256(((size_t) VG_MAX_CELLS + 1) * sizeof(int));
    v.chunk_sum = reinterpret_cast<int2*>(p);
}

// pts[0 .. U) with U = state->U on the device (U <= Umax) -> out[0 .. state->M)
int vg_run(const float* pts, int Umax, float leaf, float* out, const VgScratch& v, int sms, cudaStream_t st) {
    const float inv_leaf = 1.f / leaf;  // inverse_leaf_size_ (voxel_grid.h:236)
    vg_bbox_kernel<<<std::max(1, std::min(sms, div_up(Umax, 256))), 256, 0, st>>>(pts, v.state, inv_leaf);
    DFU_LAUNCH_OK();
    vg_zero_kernel<<<sms * 4, 256, 0, st>>>(v.state, v.count);
    DFU_LAUNCH_OK();
    vg_count_kernel<<<div_up(Umax, 256), 256, 0, st>>>(pts, v.state, inv_leaf, v.key, v.count);
    DFU_LAUNCH_OK();
    vg_chunk_sum_kernel<<<VG_MAX_CHUNKS, 256, 0, st>>>(v.state, v.count, v.chunk_sum);
    DFU_LAUNCH_OK();
    vg_chunk_scan_kernel<<<1, 1024, 0, st>>>(v.state, v.chunk_sum);
    DFU_LAUNCH_OK();
    vg_offsets_kernel<<<VG_MAX_CHUNKS, 256, 0, st>>>(v.state, v.count, v.chunk_sum, v.start, v.occ_list);
    DFU_LAUNCH_OK();
    vg_fill_kernel<<<div_up(Umax, 256), 256, 0, st>>>(v.state, v.key, v.start, v.count, v.bucket);
    DFU_LAUNCH_OK();
    vg_centroid_kernel<<<div_up(Umax, 8), 256, 0, st>>>(pts, v.state, v.start, v.occ_list, v.bucket, v.bucket2, out);
    DFU_LAUNCH_OK();
    return DFU_OK;
}

// ---- node insertion ---------------------------------------------------------------------------------------------
__global__ void append_nodes_kernel(const float* __restrict__ cent, int M, float w_new, float* __restrict__ pos_tail,
                                    float* __restrict__ w_tail) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    pos_tail[3 * (size_t) i] = cent[3 * (size_t) i];
    pos_tail[3 * (size_t) i + 1] = cent[3 * (size_t) i + 1];
    pos_tail[3 * (size_t) i + 2] = cent[3 * (size_t) i + 2];
    w_tail[i] = w_new;
}

// A cached brick stays valid when no new node can enter the 8-NN of any of its voxels: every voxel v of a brick with
// centre c and half diagonal R has its 8th neighbour within sqrt(d8(c)) + R, and |v - m| >= |c - m| - R.
__global__ void brick_invalidate_kernel(const float2* __restrict__ bounds, unsigned char* __restrict__ built, long first_brick,
                                        long n_bricks, int bd0, int bd1, float vx, float vy, float vz, const float* __restrict__ cent,
                                        int M) {
    const long p = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_bricks) return;
    if (!built[p]) return;
    const long b = first_brick + p;
    const int bx = (int) (b % bd0), by = (int) ((b / bd0) % bd1), bz = (int) (b / ((long) bd0 * bd1));
    const float cx = (bx * 8 + 3.5f) * vx, cy = (by * 8 + 3.5f) * vy, cz = (bz * 8 + 3.5f) * vz;
    const float R = 3.5f * sqrtf(vx * vx + vy * vy + vz * vz);
    const float reach = sqrtf(bounds[b].x) * 1.0001f + 2.f * R + 1e-4f;
    const float reach2 = reach * reach;
    for (int m = 0; m < M; ++m) {
        const float dx = cent[3 * m] - cx, dy = cent[3 * m + 1] - cy, dz = cent[3 * m + 2] - cz;
        if (!(dx * dx + dy * dy + dz * dz > reach2)) {  // also catches NaN / inf reach
            built[p] = 0;
            return;
        }
    }
}

int device_sms(int device) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    return sms;
}

}  // namespace

extern "C" {

int dfu_warpfield_unsupported(const dfu_warpfield* wf, const float* verts_xyz, int P, uint8_t* flags, dfu_stream stream) {
    DFU_REQUIRE(wf && (P == 0 || (verts_xyz && flags)), DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(wf->initialised, DFU_ERR_NOT_INIT, "warp field not initialised");
    if (P == 0) return DFU_OK;
    cudaStream_t st = as_stream(stream);
    int32_t* idx = nullptr;
    DFU_CUDA_OK(scratch_alloc((void**) &idx, (size_t) P * DFU_KNN * sizeof(int32_t), st));
    int rc = dfu_warpfield_knn(wf, verts_xyz, P, idx, nullptr, stream);
    if (rc == DFU_OK) {
        unsupported_kernel<<<div_up(P, 256), 256, 0, st>>>(verts_xyz, P, idx, wf->pos_w, flags);
        ++g_dfu_launches;
        if (cudaGetLastError() != cudaSuccess) rc = DFU_ERR_CUDA;
    }
    cudaFreeAsync(idx, st);
    return rc;
}

int dfu_voxel_grid_filter(const float* pts_xyz, int U, float leaf, float* out_xyz, int* M_host, dfu_stream stream) {
    DFU_REQUIRE(M_host && (U == 0 || (pts_xyz && out_xyz)), DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(U >= 0 && leaf > 0.f, DFU_ERR_INVALID, "bad size");
    *M_host = 0;
    if (U == 0) return DFU_OK;
    (void) cudaGetLastError();
    cudaStream_t st = as_stream(stream);
    int device = 0;
    DFU_CUDA_OK(cudaGetDevice(&device));
    char* mem = nullptr;
    DFU_CUDA_OK(scratch_alloc((void**) &mem, vg_scratch_bytes(U), st));
    VgScratch v;
    vg_carve(v, mem, U);
    vg_begin_kernel<<<1, 1, 0, st>>>(v.state, U);
    ++g_dfu_launches;
    int rc = vg_run(pts_xyz, U, leaf, out_xyz, v, device_sms(device), st);
    VgState h{};
    if (rc == DFU_OK && cudaMemcpyAsync(&h, v.state, sizeof(VgState), cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = DFU_ERR_CUDA;
    if (rc == DFU_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = DFU_ERR_CUDA;
    cudaFreeAsync(mem, st);
    if (rc != DFU_OK) {
        dfu_set_error("dfu_voxel_grid_filter: %s", cudaGetErrorString(cudaGetLastError()));
        return rc;
    }
    DFU_REQUIRE(!h.overflow, DFU_ERR_UNSUPPORTED, "voxel grid needs more than 2^21 cells (leaf too small for the extent)");
    *M_host = h.M;
    return DFU_OK;
}

int dfu_warpfield_cache_stats(const dfu_warpfield* wf, long long* pool_bricks_host, long long* built_bricks_host, dfu_stream stream) {
    DFU_REQUIRE(wf && pool_bricks_host && built_bricks_host, DFU_ERR_INVALID, "NULL argument");
    const BrickTable& bt = wf->bricks;
    *pool_bricks_host = (long long) bt.pool_bricks;
    *built_bricks_host = 0;
    if (!bt.built || bt.pool_bricks == 0 || bt.cache_epoch != wf->node_epoch) return DFU_OK;
    std::string host(bt.pool_bricks, '\0');
    DFU_CUDA_OK(cudaMemcpyAsync(&host[0], bt.built, bt.pool_bricks, cudaMemcpyDeviceToHost, as_stream(stream)));
    DFU_CUDA_OK(cudaStreamSynchronize(as_stream(stream)));
    long long n = 0;
    for (char c : host) n += c != 0;
    *built_bricks_host = n;
    return DFU_OK;
}

int dfu_warpfield_update(dfu_warpfield* wf, const float* verts_xyz, int P, int blend_mode, int* num_unsupported_host,
                         int* num_new_host, dfu_stream stream) {
    DFU_REQUIRE(wf && (P == 0 || verts_xyz), DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(wf->initialised, DFU_ERR_NOT_INIT, "warp field not initialised");
    DFU_REQUIRE(P >= 0, DFU_ERR_INVALID, "bad size");
    if (num_unsupported_host) *num_unsupported_host = 0;
    if (num_new_host) *num_new_host = 0;
    if (P == 0) return DFU_OK;
    DFU_GUARD(wf->device);
    cudaStream_t st = as_stream(stream);
    const int sms = device_sms(wf->device);
    const int chunks = div_up(P, 32);
    // scratch: [flags | chunk offsets | unsupported xyz | centroids xyz | voxel-grid scratch]
    const size_t o_off = up256((size_t) P), o_uns = o_off + up256((size_t) chunks * sizeof(int)),
                 o_cent = o_uns + up256((size_t) P * 3 * sizeof(float)), o_vg = o_cent + up256((size_t) P * 3 * sizeof(float));
    char* mem = nullptr;
    DFU_CUDA_OK(scratch_alloc((void**) &mem, o_vg + vg_scratch_bytes(P), st));
    unsigned char* flags = reinterpret_cast<unsigned char*>(mem);
    int* offsets = reinterpret_cast<int*>(mem + o_off);
    float* uns = reinterpret_cast<float*>(mem + o_uns);
    float* cent = reinterpret_cast<float*>(mem + o_cent);
    VgScratch v;
    vg_carve(v, mem + o_vg, P);
    VgState h{};
    int rc = dfu_warpfield_unsupported(wf, verts_xyz, P, flags, stream);
    auto launched = [&]() {
        ++g_dfu_launches;
        if (rc == DFU_OK && cudaGetLastError() != cudaSuccess) rc = DFU_ERR_CUDA;
    };
    if (rc == DFU_OK) {
        vg_begin_kernel<<<1, 1, 0, st>>>(v.state, -1);
        launched();
        flag_count_kernel<<<div_up(chunks, 8), 256, 0, st>>>(flags, P, offsets);
        launched();
        small_scan_kernel<<<1, 1024, 0, st>>>(offsets, chunks, &v.state->U);
        launched();
        flag_emit_kernel<<<div_up(chunks, 8), 256, 0, st>>>(flags, verts_xyz, P, offsets, uns);
        launched();
    }
    if (rc == DFU_OK) rc = vg_run(uns, P, 0.05f, cent, v, sms, st);  // sampler.setLeafSize(0.05, 0.05f, 0.05f)
    if (rc == DFU_OK && cudaMemcpyAsync(&h, v.state, sizeof(VgState), cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = DFU_ERR_CUDA;
    if (rc == DFU_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = DFU_ERR_CUDA;
    if (rc == DFU_OK && h.overflow) {
        dfu_set_error("dfu_warpfield_update: the unsupported vertices span more than 2^21 cells of 5 cm");
        rc = DFU_ERR_UNSUPPORTED;
    }
    // U was zeroed on overflow; otherwise it is the number of unsupported vertices
    const int M = rc == DFU_OK ? h.M : 0;
    if (rc == DFU_OK && num_unsupported_host) *num_unsupported_host = h.U;
    if (rc == DFU_OK && M > 0) {
        const int N = wf->N, N2 = N + M;
        float* nodes = nullptr;  // [dq N2*8 | pos N2*3 | w N2]: the tail of dq stays 16-byte aligned for the blend
        if (scratch_alloc((void**) &nodes, (size_t) N2 * 12 * sizeof(float), st) != cudaSuccess) rc = DFU_ERR_CUDA;
        if (rc == DFU_OK) {
            float *dq = nodes, *pos = nodes + (size_t) N2 * 8, *w = nodes + (size_t) N2 * 11;
            rc = dfu_warpfield_get_nodes(wf, pos, dq, w, stream);
            // dg_se3 = calcDQB(dg_v) against the old nodes (warp_field.cpp:79), dg_w = 2 * epsilon (:80)
            if (rc == DFU_OK) rc = dfu_warpfield_blend(wf, cent, M, dq + (size_t) N * 8, blend_mode, stream);
            if (rc == DFU_OK) {
                append_nodes_kernel<<<div_up(M, 256), 256, 0, st>>>(cent, M, 2.f * wf->epsilon, pos + (size_t) N * 3, w + N);
                launched();
            }
            // voxel cache: drop only the bricks a new node can reach (old node ids and weights are unchanged elsewhere)
            BrickTable& bt = wf->bricks;
            bool keep_cache = false;
            if (rc == DFU_OK && bt.valid && bt.built && bt.pool_bricks > 0 && bt.cache_epoch == wf->node_epoch &&
                bt.node_epoch == wf->node_epoch && M <= 4096 && N2 <= 65535) {
                const int bd0 = bt.dims[0] / 8, bd1 = bt.dims[1] / 8;
                const long first = (long) bd0 * bd1 * bt.pool_zb0;
                brick_invalidate_kernel<<<div_up((long) bt.pool_bricks, 256), 256, 0, st>>>(
                    bt.bounds, bt.built, first, (long) bt.pool_bricks, bd0, bd1, bt.voxel[0], bt.voxel[1], bt.voxel[2], cent, M);
                launched();
                keep_cache = rc == DFU_OK;
            }
            if (rc == DFU_OK) rc = dfu_warpfield_init(wf, wf->epsilon, pos, dq, w, N2, stream);  // re-pack + node grid (:86-94)
            if (rc == DFU_OK && keep_cache) bt.cache_epoch = wf->node_epoch;
            cudaFreeAsync(nodes, st);
        }
        if (rc == DFU_OK && num_new_host) *num_new_host = M;
    }
    cudaFreeAsync(mem, st);
    if (rc == DFU_ERR_CUDA) dfu_set_error("dfu_warpfield_update: CUDA error %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

}  // extern "C"
