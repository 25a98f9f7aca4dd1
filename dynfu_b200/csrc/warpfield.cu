// Warp field on the device: node storage, exact 8-NN, blending, point warping.
// Replaces class Warpfield (include/dynfu/warp_field.hpp:32-78, src/dynfu/warp_field.cpp), Node
// (src/dynfu/utils/node.cpp) and the DualQuaternion arithmetic they call, for whole arrays of points.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "dfu_internal.h"
#include "blend.cuh"
#include "knn.cuh"

using namespace dfu;

// ------------------------------------------------------------------------------------------------
// error plumbing
static thread_local char g_err[512] = "";
void dfu_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* dfu_last_error(void) { return g_err; }
extern "C" int dfu_version(void) { return DFU_VERSION; }
unsigned long long g_dfu_launches = 0;
extern "C" unsigned long long dfu_launch_count(void) { return g_dfu_launches; }

extern "C" int dfu_device_check(int device) {
    int n = 0;
    DFU_CUDA_OK(cudaGetDeviceCount(&n));
    DFU_REQUIRE(device >= 0 && device < n, DFU_ERR_CUDA, "no such CUDA device");
    cudaDeviceProp p;
    DFU_CUDA_OK(cudaGetDeviceProperties(&p, device));
    DFU_REQUIRE(p.major == 10, DFU_ERR_CUDA, "device is not sm_100 (this library ships sm_100a code only)");
    return DFU_OK;
}

namespace {


// ---- node packing ------------------------------------------------------------------------------
__global__ void pack_nodes_kernel(const float* __restrict__ pos, const float* __restrict__ dq,
                                  const float* __restrict__ dgw, int N, int Npad, float4* __restrict__ pos_w,
                                  float4* __restrict__ real, float4* __restrict__ dual) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Npad) return;
    if (i < N) {
        if (pos) pos_w[i] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], dgw[i]);
        if (dq) {
            real[i] = make_float4(dq[8 * i], dq[8 * i + 1], dq[8 * i + 2], dq[8 * i + 3]);
            dual[i] = make_float4(dq[8 * i + 4], dq[8 * i + 5], dq[8 * i + 6], dq[8 * i + 7]);
        }
    } else {
        if (pos) pos_w[i] = make_float4(INFINITY, INFINITY, INFINITY, 1.f);
        if (dq) {
            real[i] = make_float4(1.f, 0.f, 0.f, 0.f);
            dual[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

__global__ void unpack_nodes_kernel(const float4* __restrict__ pos_w, const float4* __restrict__ real,
                                    const float4* __restrict__ dual, int N, float* __restrict__ pos,
                                    float* __restrict__ dq, float* __restrict__ dgw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float4 p = pos_w[i];
    if (pos) {
        pos[3 * i] = p.x; pos[3 * i + 1] = p.y; pos[3 * i + 2] = p.z;
    }
    if (dgw) dgw[i] = p.w;
    if (dq) {
        const float4 r = real[i], d = dual[i];
        dq[8 * i] = r.x; dq[8 * i + 1] = r.y; dq[8 * i + 2] = r.z; dq[8 * i + 3] = r.w;
        dq[8 * i + 4] = d.x; dq[8 * i + 5] = d.y; dq[8 * i + 6] = d.z; dq[8 * i + 7] = d.w;
    }
}

// flags[0] = 1 iff every node is a pure translation: real == (1,0,0,0) and dual.w == 0 (the only state the
// reference ever produces, src/dynfu/utils/opt_solver.cpp:280-281); flags[1] = max dg_w (float bits);
// flags[2] = max over nodes of max(|dual.x|,|dual.y|,|dual.z|) (float bits)
__global__ void node_flags_kernel(const float4* __restrict__ pos_w, const float4* __restrict__ real,
                                  const float4* __restrict__ dual, int N, int* __restrict__ flags) {
    __shared__ int s_ok;
    __shared__ unsigned s_maxw, s_maxd;
    if (threadIdx.x == 0) {
        s_ok = 1;
        s_maxw = 0u;
        s_maxd = 0u;
    }
    __syncthreads();
    int ok = 1;
    float mw = 0.f, md = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float4 r = real[i], d = dual[i];
        ok &= (r.x == 1.f && r.y == 0.f && r.z == 0.f && r.w == 0.f && d.x == 0.f);
        mw = fmaxf(mw, pos_w[i].w);
        md = fmaxf(md, fmaxf(fabsf(d.y), fmaxf(fabsf(d.z), fabsf(d.w))));
    }
    if (!ok) atomicAnd(&s_ok, 0);
    atomicMax(&s_maxw, __float_as_uint(mw));  // non-negative floats: uint order == float order
    atomicMax(&s_maxd, __float_as_uint(md));
    __syncthreads();
    if (threadIdx.x == 0) {
        flags[0] = s_ok;
        flags[1] = (int) s_maxw;
        flags[2] = (int) s_maxd;
    }
}

// dg_se3 := DQ(0,0,0,t) * dg_se3   (src/dynfu/utils/node.cpp:19-23, opt_solver.cpp:278-283)
__global__ void update_translations_kernel(float4* __restrict__ real, float4* __restrict__ dual,
                                           const float* __restrict__ t, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const DQ inc = dq_from_translation(t[3 * i], t[3 * i + 1], t[3 * i + 2]);
    const DQ old{make_quat(real[i]), make_quat(dual[i])};
    const DQ res = dq_mul(inc, old);
    real[i] = to_float4(res.real);
    dual[i] = to_float4(res.dual);
}

// ---- the per-point kernel: kNN, optionally followed by blend / warp -------------------------------
enum { OP_KNN = 0, OP_BLEND = 1, OP_WARP = 2, OP_BOUNDS = 3, OP_GRAPH = 4, OP_NODEGRAPH = 5 };

struct PointArgs {
    const float4 *pos_w, *real, *dual;
    int Npad;
    const float* q;   // Q*3 query points (OP_BOUNDS: unused)
    int Q;
    int32_t* idx;     // OP_KNN
    float* dist2;     // OP_KNN (may be null)
    float* dq_out;    // OP_BLEND
    const float* n_in;  // OP_WARP
    float* v_out;
    float* n_out;
    int blend_mode, normal_mode;
    // OP_GRAPH: data graph of the solver (idx = neighbours, wts = node weights, dvec = live - canon)
    float* wts;
    const float* live;
    float* dvec;
    int* deg;         // OP_GRAPH (may be null): per-node count of referencing points, incremented here
    int32_t* rank;    // OP_GRAPH (may be null, needs deg): per edge, how many references its node had before this one
    // OP_BOUNDS: queries are the centres of the 8x8x8 bricks of a volume
    float2* bounds;
    int bdim[3];
    float voxel[3];
};

// G = 8 consecutive lanes cooperate on one query: split scan + shuffle merge (knn.cuh), then lane `sub`
// evaluates the (FP64) weight of neighbour `sub` -- 8 lanes, 8 neighbours -- and the lanes share them by shuffle.
constexpr int QG = 8;                  // lanes per query
constexpr int QPB = 256 / QG;          // queries per CTA

DFU_DEV float pick(const float (&a)[DFU_KNN], int k) {
    float r = a[0];
#pragma unroll
    for (int i = 1; i < DFU_KNN; ++i) r = (i == k) ? a[i] : r;
    return r;
}
DFU_DEV int pick(const int (&a)[DFU_KNN], int k) {
    int r = a[0];
#pragma unroll
    for (int i = 1; i < DFU_KNN; ++i) r = (i == k) ? a[i] : r;
    return r;
}

template <int OP>
__global__ void __launch_bounds__(256) points_kernel(const PointArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    KnnSmem& sm = *reinterpret_cast<KnnSmem*>(smem_raw);
    const int sub = threadIdx.x & (QG - 1);
    const long q = (long) blockIdx.x * QPB + (threadIdx.x / QG);
    const bool active = q < a.Q;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        if (OP == OP_BOUNDS) {
            const int bx = (int) (q % a.bdim[0]), by = (int) ((q / a.bdim[0]) % a.bdim[1]), bz = (int) (q / ((long) a.bdim[0] * a.bdim[1]));
            qx = (bx * 8 + 3.5f) * a.voxel[0];
            qy = (by * 8 + 3.5f) * a.voxel[1];
            qz = (bz * 8 + 3.5f) * a.voxel[2];
        } else if (OP == OP_NODEGRAPH) {  // queries are the node positions themselves
            const float4 p = a.pos_w[q];
            qx = p.x; qy = p.y; qz = p.z;
        } else {
            qx = a.q[3 * (size_t) q];
            qy = a.q[3 * (size_t) q + 1];
            qz = a.q[3 * (size_t) q + 2];
        }
    }
    V3 nn{0.f, 0.f, 0.f};  // read before anything is written: v_out / n_out may alias the inputs
    if (OP == OP_WARP && active && a.n_in) nn = V3{a.n_in[3 * (size_t) q], a.n_in[3 * (size_t) q + 1], a.n_in[3 * (size_t) q + 2]};
    Top8 t;
    knn8_scan_block_split<QG>(sm, a.pos_w, a.Npad, qx, qy, qz, active, sub, t);
    // a warp holds 4 queries; inactive queries only exist in the last CTA and come in whole groups of 8 lanes,
    // so the full-mask shuffles below are executed by every lane of the warp
    const int my_i = pick(t.i, sub);
    const float my_d = pick(t.d, sub);

    if (OP == OP_KNN || OP == OP_NODEGRAPH) {
        if (active) {
            a.idx[(size_t) q * DFU_KNN + sub] = my_i;
            if (a.dist2) a.dist2[(size_t) q * DFU_KNN + sub] = my_d;
        }
        return;
    }
    if (OP == OP_BOUNDS) {
        if (active && sub == 0) a.bounds[q] = make_float2(t.d[DFU_KNN - 1], t.d[0]);
        return;
    }
    // weight of neighbour `sub` (Node::getTransformationWeight)
    float my_w = 0.f;
    if (active && my_i >= 0) {
        const float4 nd = __ldg(&a.pos_w[my_i]);
        my_w = node_weight(nd.x, nd.y, nd.z, nd.w, qx, qy, qz, my_d);
    }
    if (OP == OP_GRAPH) {
        // CombinedSolver::initializeDataGraph (src/dynfu/utils/opt_solver.cpp:56-72) + the per-edge weight of
        // energy.t:15-17,49-52 + (live - canon) of energy.t:55
        if (active) {
            a.idx[(size_t) q * DFU_KNN + sub] = my_i;
            a.wts[(size_t) q * DFU_KNN + sub] = my_w;
            if (a.deg && my_i >= 0) {
                const int rk = atomicAdd(&a.deg[my_i], 1);
                if (a.rank) a.rank[(size_t) q * DFU_KNN + sub] = rk;
            }
            if (sub < 3 && a.dvec) a.dvec[3 * (size_t) q + sub] = a.live[3 * (size_t) q + sub] - (sub == 0 ? qx : (sub == 1 ? qy : qz));
        }
        return;
    }
    float w[DFU_KNN];
    const int base = (threadIdx.x & 31) & ~(QG - 1);
#pragma unroll
    for (int k = 0; k < DFU_KNN; ++k) w[k] = __shfl_sync(0xffffffffu, my_w, base + k);
    if (!active) return;
    const DQ b = blend(a.blend_mode, t, w, a.real, a.dual);  // all 8 lanes compute the same chain
    if (OP == OP_BLEND) {
        const float o[8] = {b.real.w, b.real.x, b.real.y, b.real.z, b.dual.w, b.dual.x, b.dual.y, b.dual.z};
        a.dq_out[(size_t) q * 8 + sub] = pick(o, sub);
    } else {
        // Warpfield::warpToLive (src/dynfu/warp_field.cpp:150-171)
        const V3 r = dq_transform_vertex(b, V3{qx, qy, qz});
        if (sub < 3) a.v_out[3 * (size_t) q + sub] = sub == 0 ? r.x : (sub == 1 ? r.y : r.z);
        if (a.n_in && a.n_out) {
            const V3 rn = a.normal_mode == DFU_NORMAL_REF ? dq_transform_vertex(b, nn) : dq_rotate(b, nn);
            if (sub >= 3 && sub < 6) a.n_out[3 * (size_t) q + sub - 3] = sub == 3 ? rn.x : (sub == 4 ? rn.y : rn.z);
        }
    }
}

// =====================================================================================================
// Grid-bucketed exact kNN for point queries (the north-star's "(1) grid-bucketed ... kNN").
//
// Nodes are binned into a uniform grid (cell edge ~2.5 mean nearest-neighbour distances).  A query visits the
// cells in shells of growing Chebyshev radius around its own cell and stops as soon as its current 8th
// distance is smaller than the distance to everything not yet visited, so it evaluates tens of candidates
// instead of all N.  Distances use the same expression as everywhere else (bit-exact) and the top-8 is kept
// by the (dist2, index) key explicitly, because cells are not visited in index order.

// bbox, mean nearest-neighbour distance (sampled) -> grid descriptor.  One CTA of 1024 threads.
__global__ void __launch_bounds__(1024) grid_desc_kernel(const float4* __restrict__ pos_w, int N, GridDesc* __restrict__ g) {
    __shared__ float s_min[3][32], s_max[3][32], s_nn[32];
    __shared__ int s_cnt[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < N; i += 1024) {
        const float4 p = pos_w[i];
        mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
        mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
    }
    // nearest-neighbour distance of up to 1024 sample nodes
    float nn = 0.f;
    int cnt = 0;
    const int stride = max(1, N / 1024);
    const int i0 = tid * stride;
    if (i0 < N && N > 1) {
        const float4 q = pos_w[i0];
        float best = INFINITY;
        for (int j = 0; j < N; ++j) {
            if (j == i0) continue;
            const float4 p = pos_w[j];
            const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
            best = fminf(best, dx * dx + dy * dy + dz * dz);
        }
        nn = sqrtf(best);
        cnt = 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        for (int c = 0; c < 3; ++c) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
        nn += __shfl_xor_sync(0xffffffffu, nn, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) {
        for (int c = 0; c < 3; ++c) {
            s_min[c][wid] = mn[c];
            s_max[c][wid] = mx[c];
        }
        s_nn[wid] = nn;
        s_cnt[wid] = cnt;
    }
    __syncthreads();
    if (tid == 0) {
        float nnsum = 0.f;
        int c_ = 0;
        for (int w = 0; w < 32; ++w) {
            for (int c = 0; c < 3; ++c) {
                mn[c] = fminf(mn[c], s_min[c][w]);
                mx[c] = fmaxf(mx[c], s_max[c][w]);
            }
            nnsum += s_nn[w];
            c_ += s_cnt[w];
        }
        const float ext = fmaxf(mx[0] - mn[0], fmaxf(mx[1] - mn[1], mx[2] - mn[2]));
        float h = c_ > 0 ? 2.5f * nnsum / (float) c_ : 1.f;
        h = fmaxf(h, ext / (float) (DFU_GRID_MAX_DIM - 1));
        h = fmaxf(h, 1e-6f);
        g->ox = mn[0]; g->oy = mn[1]; g->oz = mn[2];
        g->h = h;
        g->inv_h = 1.f / h;
        g->nx = min(DFU_GRID_MAX_DIM, (int) ((mx[0] - mn[0]) * g->inv_h) + 1);
        g->ny = min(DFU_GRID_MAX_DIM, (int) ((mx[1] - mn[1]) * g->inv_h) + 1);
        g->nz = min(DFU_GRID_MAX_DIM, (int) ((mx[2] - mn[2]) * g->inv_h) + 1);
        g->n_occ = 0;
    }
}

DFU_DEV int grid_coord(float v, float o, float inv_h, int n) { return min(n - 1, max(0, (int) floorf((v - o) * inv_h))); }
DFU_DEV int grid_cell_of(const GridDesc& g, float x, float y, float z) {
    return grid_coord(x, g.ox, g.inv_h, g.nx) + g.nx * (grid_coord(y, g.oy, g.inv_h, g.ny) + g.ny * grid_coord(z, g.oz, g.inv_h, g.nz));
}
__global__ void grid_count_kernel(const float4* __restrict__ pos_w, int N, const GridDesc* __restrict__ gd, int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const GridDesc g = *gd;
    const float4 p = pos_w[i];
    atomicAdd(&count[grid_cell_of(g, p.x, p.y, p.z)], 1);
}
// exclusive scan of count[0..ncells) -> start[0..ncells]; single CTA, ncells read from the descriptor
__global__ void __launch_bounds__(1024) grid_scan_kernel(const int* __restrict__ count, const GridDesc* __restrict__ gd, int* __restrict__ start) {
    __shared__ int sh[1024];
    const int n = gd->nx * gd->ny * gd->nz;
    const int per = (n + 1023) / 1024;
    const int lo = min(n, (int) threadIdx.x * per), hi = min(n, lo + per);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += count[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = (int) threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += v;
        __syncthreads();
    }
    int run = sh[threadIdx.x] - s;
    for (int i = lo; i < hi; ++i) {
        start[i] = run;
        run += count[i];
    }
    if (threadIdx.x == 1023) start[n] = sh[1023];
}
__global__ void grid_fill_kernel(const float4* __restrict__ pos_w, int N, const GridDesc* __restrict__ gd, const int* __restrict__ start,
                                 int* __restrict__ count, float4* __restrict__ sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const GridDesc g = *gd;
    const float4 p = pos_w[i];
    const int c = grid_cell_of(g, p.x, p.y, p.z);
    const int slot = atomicSub(&count[c], 1) - 1;  // order inside a cell is irrelevant: the search uses the (dist2, idx) key
    sorted[start[c] + slot] = make_float4(p.x, p.y, p.z, __int_as_float(i));
}

// list of the non-empty cells (order arbitrary: it only affects the visiting order, never the result)
__global__ void grid_occ_kernel(const int* __restrict__ start, GridDesc* __restrict__ gd, int* __restrict__ occ) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = gd->nx * gd->ny * gd->nz;
    if (c < n && start[c + 1] > start[c]) occ[atomicAdd(&gd->n_occ, 1)] = c;
}

// ---- exact 8-NN through the grid ------------------------------------------------------------------------------
// The search key (dist2, idx) is packed into one 64-bit word, distance bits high: squared distances are >= 0, so their
// IEEE bit patterns order like unsigned integers and ONE unsigned compare is the lexicographic compare.  The list is
// kept ascending in registers; an insertion is 8 compare-selects.
struct Keys8 {
    unsigned long long k[DFU_KNN];
};
DFU_DEV unsigned long long knn_key(float d, int idx) { return ((unsigned long long) __float_as_uint(d) << 32) | (unsigned) idx; }
DFU_DEV void keys8_insert(Keys8& t, unsigned long long key) {
#pragma unroll
    for (int k = DFU_KNN - 1; k > 0; --k) {
        const bool shift = key < t.k[k - 1];
        const bool here = !shift && key < t.k[k];
        t.k[k] = shift ? t.k[k - 1] : (here ? key : t.k[k]);
    }
    if (key < t.k[0]) t.k[0] = key;
}
// all nodes of the cells [c0, c1] of one grid row: consecutive cells are consecutive in `sorted`
DFU_DEV void grid_visit_range(const int* __restrict__ start, const float4* __restrict__ sorted, int c0, int c1, float qx, float qy,
                              float qz, Keys8& t) {
    const int lo = __ldg(&start[c0]), hi = __ldg(&start[c1 + 1]);
    for (int j = lo; j < hi; ++j) {
        const float4 p = __ldg(&sorted[j]);
        const unsigned long long key = knn_key(dist2(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w));
        if (key < t.k[DFU_KNN - 1]) keys8_insert(t, key);
    }
}

// Shells of growing Chebyshev radius around the query's cell (a query near the nodes is done after the first 3x3x3
// block); once the next shell would cost more cell visits than there are NON-EMPTY cells, a query that is still not
// settled sweeps the list of non-empty cells instead, pruning each by its box distance, so the cost stays bounded
// however far the query is from the nodes.
// radius 0 and 1 around the query's cell: up to nine row segments (cells of one grid row are contiguous in `sorted`);
// segment s of the block, s < grid_block_segments()
DFU_DEV int grid_block_segments(const GridDesc& g, int cy, int cz) {
    return (min(g.nz - 1, cz + 1) - max(0, cz - 1) + 1) * (min(g.ny - 1, cy + 1) - max(0, cy - 1) + 1);
}
DFU_DEV void grid_visit_block_segment(const GridDesc& g, const int* __restrict__ start, const float4* __restrict__ sorted, int cx,
                                      int cy, int cz, int s, float qx, float qy, float qz, Keys8& t) {
    const int y0 = max(0, cy - 1), ny = min(g.ny - 1, cy + 1) - y0 + 1;
    const int z = max(0, cz - 1) + s / ny, y = y0 + s % ny;
    const int row = g.nx * (y + g.ny * z);
    grid_visit_range(start, sorted, row + max(0, cx - 1), row + min(g.nx - 1, cx + 1), qx, qy, qz, t);
}

// the rest of the search once the 3x3x3 block has been visited: settle test, further shells, sweep of the non-empty cells
DFU_DEV void knn8_grid_continue(const GridDesc& g, const int* __restrict__ start, const float4* __restrict__ sorted,
                                const int* __restrict__ occ, int cx, int cy, int cz, float qx, float qy, float qz, Keys8& t) {
    bool settled = false;
    int r = 1, rd = -1;  // rd: every cell within this Chebyshev radius has been visited
    for (;; ++r) {
        if (r > 1) {
            if ((2 * r + 1) * (2 * r + 1) * (2 * r + 1) > 2 * g.n_occ + 27) break;
            const int z0 = max(0, cz - r), z1 = min(g.nz - 1, cz + r);
            const int y0 = max(0, cy - r), y1 = min(g.ny - 1, cy + r);
            const int x0 = max(0, cx - r), x1 = min(g.nx - 1, cx + r);
            for (int z = z0; z <= z1; ++z)
                for (int y = y0; y <= y1; ++y) {
                    const bool face = (abs(z - cz) == r) || (abs(y - cy) == r);
                    const int row = g.nx * (y + g.ny * z);
                    if (face) {
                        grid_visit_range(start, sorted, row + x0, row + x1, qx, qy, qz, t);
                    } else {  // interior rows of the shell: only the two end cells
                        if (cx - r >= 0) grid_visit_range(start, sorted, row + cx - r, row + cx - r, qx, qy, qz, t);
                        if (cx + r < g.nx) grid_visit_range(start, sorted, row + cx + r, row + cx + r, qx, qy, qz, t);
                    }
                }
        }
        rd = r;
        // distance from the query to everything not visited yet: the nearest face of the visited box that
        // still has cells beyond it
        float L = INFINITY;
        if (cx - r > 0) L = fminf(L, qx - (g.ox + (float) (cx - r) * g.h));
        if (cx + r < g.nx - 1) L = fminf(L, (g.ox + (float) (cx + r + 1) * g.h) - qx);
        if (cy - r > 0) L = fminf(L, qy - (g.oy + (float) (cy - r) * g.h));
        if (cy + r < g.ny - 1) L = fminf(L, (g.oy + (float) (cy + r + 1) * g.h) - qy);
        if (cz - r > 0) L = fminf(L, qz - (g.oz + (float) (cz - r) * g.h));
        if (cz + r < g.nz - 1) L = fminf(L, (g.oz + (float) (cz + r + 1) * g.h) - qz);
        // margin: binning and face coordinates are rounded (coordinates are O(1) m: 1e-5 m is > 100 ulp)
        const float Ls = L - 1e-5f - 1e-5f * g.h;
        const float d8 = __uint_as_float((unsigned) (t.k[DFU_KNN - 1] >> 32));
        if (L == INFINITY || (Ls > 0.f && d8 < Ls * Ls)) {
            settled = true;
            break;
        }
    }
    if (!settled) {
        const int n_occ = g.n_occ;
        for (int e = 0; e < n_occ; ++e) {
            const int c = __ldg(&occ[e]);
            const int ix = c % g.nx, iy = (c / g.nx) % g.ny, iz = c / (g.nx * g.ny);
            if (abs(ix - cx) <= rd && abs(iy - cy) <= rd && abs(iz - cz) <= rd) continue;
            const float lx = g.ox + (float) ix * g.h, ly = g.oy + (float) iy * g.h, lz = g.oz + (float) iz * g.h;
            const float ddx = fmaxf(0.f, fmaxf(lx - qx, qx - (lx + g.h))), ddy = fmaxf(0.f, fmaxf(ly - qy, qy - (ly + g.h))),
                        ddz = fmaxf(0.f, fmaxf(lz - qz, qz - (lz + g.h)));
            const float dl = fmaxf(0.f, sqrtf(ddx * ddx + ddy * ddy + ddz * ddz) - 1e-5f - 1e-5f * g.h);
            if (dl * dl <= __uint_as_float((unsigned) (t.k[DFU_KNN - 1] >> 32))) grid_visit_range(start, sorted, c, c, qx, qy, qz, t);
        }
    }
}
DFU_DEV void keys8_to_top8(const Keys8& t, Top8& out) {
#pragma unroll
    for (int k = 0; k < DFU_KNN; ++k) {
        out.d[k] = __uint_as_float((unsigned) (t.k[k] >> 32));
        const int idx = (int) (unsigned) (t.k[k] & 0xffffffffull);
        out.i[k] = idx == 0x7fffffff ? -1 : idx;
    }
}

// Shells of growing Chebyshev radius around the query's cell (a query near the nodes is done after the first 3x3x3
// block); once the next shell would cost more cell visits than there are NON-EMPTY cells, a query that is still not
// settled sweeps the list of non-empty cells instead, pruning each by its box distance, so the cost stays bounded
// however far the query is from the nodes.
DFU_DEV void knn8_grid(const GridDesc& g, const int* __restrict__ start, const float4* __restrict__ sorted,
                       const int* __restrict__ occ, float qx, float qy, float qz, Top8& out) {
    Keys8 t;
    const unsigned long long none = knn_key(INFINITY, 0x7fffffff);  // "no node": loses every tie
#pragma unroll
    for (int k = 0; k < DFU_KNN; ++k) t.k[k] = none;
    const int cx = grid_coord(qx, g.ox, g.inv_h, g.nx), cy = grid_coord(qy, g.oy, g.inv_h, g.ny), cz = grid_coord(qz, g.oz, g.inv_h, g.nz);
    const int nseg = grid_block_segments(g, cy, cz);
    for (int s = 0; s < nseg; ++s) grid_visit_block_segment(g, start, sorted, cx, cy, cz, s, qx, qy, qz, t);
    knn8_grid_continue(g, start, sorted, occ, cx, cy, cz, qx, qy, qz, t);
    keys8_to_top8(t, out);
}

// The same search with G lanes per query (the data-graph build of the solver, once per frame over all surface points): one
// thread per query is a chain of ~40 dependent candidate visits and 8 FP64 weight evaluations, and 75 k queries fill a quarter
// of the machine's thread slots.  The lanes of a group take the row segments of the 3x3x3 block in turn, merge their sorted
// lists with a shuffle butterfly (the 8 smallest of two ascending lists form a bitonic sequence), run the settle test together
// (further shells are rare and done redundantly), and share the 8 weight evaluations.  Same keys, same result, bit for bit.
// Measured on the bench scene (initializeProblemInstance, B200): G = 1: 0.113-0.119 ms, 2: 0.105-0.115, 4: 0.110-0.118,
// 8: 0.126-0.134 (the merge network outweighs the shorter chains) -- the default is 2; DFU_GRAPH_LANES=1|2|4 selects.
template <int G>
DFU_DEV void keys8_merge_group(Keys8& t) {
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
        unsigned long long o[DFU_KNN];
#pragma unroll
        for (int k = 0; k < DFU_KNN; ++k) o[k] = __shfl_xor_sync(0xffffffffu, t.k[DFU_KNN - 1 - k], off);  // partner's list, reversed
#pragma unroll
        for (int k = 0; k < DFU_KNN; ++k) t.k[k] = o[k] < t.k[k] ? o[k] : t.k[k];
#pragma unroll
        for (int st = DFU_KNN / 2; st > 0; st >>= 1)
#pragma unroll
            for (int k = 0; k < DFU_KNN; ++k)
                if ((k & st) == 0 && t.k[k + st] < t.k[k]) {
                    const unsigned long long tmp = t.k[k];
                    t.k[k] = t.k[k + st];
                    t.k[k + st] = tmp;
                }
    }
}

template <int G>
__global__ void __launch_bounds__(128) points_grid_graph_kernel(const PointArgs a, const GridDesc* __restrict__ gd,
                                                                const int* __restrict__ start, const float4* __restrict__ sorted,
                                                                const int* __restrict__ occ) {
    const int sub = threadIdx.x & (G - 1);
    const long qraw = ((long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const bool active = qraw < a.Q;
    const long q = active ? qraw : a.Q - 1;  // (inactive groups replay the last query: every lane takes part in the shuffles)
    const GridDesc g = *gd;
    const float qx = a.q[3 * (size_t) q], qy = a.q[3 * (size_t) q + 1], qz = a.q[3 * (size_t) q + 2];
    Keys8 t;
    const unsigned long long none = knn_key(INFINITY, 0x7fffffff);
#pragma unroll
    for (int k = 0; k < DFU_KNN; ++k) t.k[k] = none;
    const int cx = grid_coord(qx, g.ox, g.inv_h, g.nx), cy = grid_coord(qy, g.oy, g.inv_h, g.ny), cz = grid_coord(qz, g.oz, g.inv_h, g.nz);
    const int nseg = grid_block_segments(g, cy, cz);
    for (int s = sub; s < nseg; s += G) grid_visit_block_segment(g, start, sorted, cx, cy, cz, s, qx, qy, qz, t);
    keys8_merge_group<G>(t);
    knn8_grid_continue(g, start, sorted, occ, cx, cy, cz, qx, qy, qz, t);  // (uniform over the group)
    // lane `sub` owns neighbours sub, sub + G, ...
#pragma unroll
    for (int j = 0; j < DFU_KNN / G; ++j) {
        unsigned long long mine = t.k[j * G];
#pragma unroll
        for (int k = 1; k < G; ++k) mine = sub == k ? t.k[j * G + k] : mine;
        const float d = __uint_as_float((unsigned) (mine >> 32));
        const int raw = (int) (unsigned) (mine & 0xffffffffull);
        const int idx = raw == 0x7fffffff ? -1 : raw;
        float w = 0.f;
        if (idx >= 0) {
            const float4 nd = __ldg(&a.pos_w[idx]);
            w = node_weight(nd.x, nd.y, nd.z, nd.w, qx, qy, qz, d);
        }
        if (active) {
            a.idx[(size_t) q * DFU_KNN + j * G + sub] = idx;
            a.wts[(size_t) q * DFU_KNN + j * G + sub] = w;
            if (a.deg && idx >= 0) {
                const int rk = atomicAdd(&a.deg[idx], 1);
                if (a.rank) a.rank[(size_t) q * DFU_KNN + j * G + sub] = rk;
            }
        }
    }
    if (active && a.dvec) {
#pragma unroll
        for (int c = sub; c < 3; c += G) {
            const float qc = c == 0 ? qx : (c == 1 ? qy : qz);
            a.dvec[3 * (size_t) q + c] = a.live[3 * (size_t) q + c] - qc;
        }
    }
}

template <int OP>
__global__ void __launch_bounds__(128) points_grid_kernel(const PointArgs a, const GridDesc* __restrict__ gd, const int* __restrict__ start,
                                                          const float4* __restrict__ sorted, const int* __restrict__ occ) {
    const long q = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.Q) return;
    const GridDesc g = *gd;
    float qx, qy, qz;
    if (OP == OP_NODEGRAPH) {
        const float4 p = a.pos_w[q];
        qx = p.x; qy = p.y; qz = p.z;
    } else {
        qx = a.q[3 * (size_t) q]; qy = a.q[3 * (size_t) q + 1]; qz = a.q[3 * (size_t) q + 2];
    }
    V3 nn{0.f, 0.f, 0.f};
    if (OP == OP_WARP && a.n_in) nn = V3{a.n_in[3 * (size_t) q], a.n_in[3 * (size_t) q + 1], a.n_in[3 * (size_t) q + 2]};
    Top8 t;
    knn8_grid(g, start, sorted, occ, qx, qy, qz, t);
    if (OP == OP_KNN || OP == OP_NODEGRAPH) {
        int4* o = reinterpret_cast<int4*>(a.idx + (size_t) q * DFU_KNN);
        o[0] = make_int4(t.i[0], t.i[1], t.i[2], t.i[3]);
        o[1] = make_int4(t.i[4], t.i[5], t.i[6], t.i[7]);
        if (a.dist2) {
            float4* d = reinterpret_cast<float4*>(a.dist2 + (size_t) q * DFU_KNN);
            d[0] = make_float4(t.d[0], t.d[1], t.d[2], t.d[3]);
            d[1] = make_float4(t.d[4], t.d[5], t.d[6], t.d[7]);
        }
        return;
    }
    float w[DFU_KNN];
    neighbour_weights(t, qx, qy, qz, a.pos_w, w);
    if (OP == OP_GRAPH) {
        int4* o = reinterpret_cast<int4*>(a.idx + (size_t) q * DFU_KNN);
        o[0] = make_int4(t.i[0], t.i[1], t.i[2], t.i[3]);
        o[1] = make_int4(t.i[4], t.i[5], t.i[6], t.i[7]);
        float4* ww = reinterpret_cast<float4*>(a.wts + (size_t) q * DFU_KNN);
        ww[0] = make_float4(w[0], w[1], w[2], w[3]);
        ww[1] = make_float4(w[4], w[5], w[6], w[7]);
        if (a.dvec) {
            a.dvec[3 * (size_t) q] = a.live[3 * (size_t) q] - qx;
            a.dvec[3 * (size_t) q + 1] = a.live[3 * (size_t) q + 1] - qy;
            a.dvec[3 * (size_t) q + 2] = a.live[3 * (size_t) q + 2] - qz;
        }
        if (a.deg) {
#pragma unroll
            for (int k = 0; k < DFU_KNN; ++k)
                if (t.i[k] >= 0) {
                    const int rk = atomicAdd(&a.deg[t.i[k]], 1);
                    if (a.rank) a.rank[(size_t) q * DFU_KNN + k] = rk;
                }
        }
        return;
    }
    const DQ b = blend(a.blend_mode, t, w, a.real, a.dual);
    if (OP == OP_BLEND) {
        float4* o = reinterpret_cast<float4*>(a.dq_out + (size_t) q * 8);
        o[0] = make_float4(b.real.w, b.real.x, b.real.y, b.real.z);
        o[1] = make_float4(b.dual.w, b.dual.x, b.dual.y, b.dual.z);
    } else {
        const V3 r = dq_transform_vertex(b, V3{qx, qy, qz});
        a.v_out[3 * (size_t) q] = r.x; a.v_out[3 * (size_t) q + 1] = r.y; a.v_out[3 * (size_t) q + 2] = r.z;
        if (a.n_in && a.n_out) {
            const V3 rn = a.normal_mode == DFU_NORMAL_REF ? dq_transform_vertex(b, nn) : dq_rotate(b, nn);
            a.n_out[3 * (size_t) q] = rn.x; a.n_out[3 * (size_t) q + 1] = rn.y; a.n_out[3 * (size_t) q + 2] = rn.z;
        }
    }
}

// Warp of a point set whose 8 nearest nodes and weights are cached (dfu_warpfield_warp_cached): same blend and transform
// arithmetic as points_*_kernel<OP_WARP>, fed from the cache instead of a search
__global__ void __launch_bounds__(256) warp_cached_kernel(const PointArgs a, const int32_t* __restrict__ ids, const float* __restrict__ wts) {
    const long q = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.Q) return;
    const float qx = a.q[3 * (size_t) q], qy = a.q[3 * (size_t) q + 1], qz = a.q[3 * (size_t) q + 2];
    V3 nn{0.f, 0.f, 0.f};
    if (a.n_in) nn = V3{a.n_in[3 * (size_t) q], a.n_in[3 * (size_t) q + 1], a.n_in[3 * (size_t) q + 2]};
    const int4 i0 = *(reinterpret_cast<const int4*>(ids) + 2 * (size_t) q), i1 = *(reinterpret_cast<const int4*>(ids) + 2 * (size_t) q + 1);
    const float4 w0 = *(reinterpret_cast<const float4*>(wts) + 2 * (size_t) q), w1 = *(reinterpret_cast<const float4*>(wts) + 2 * (size_t) q + 1);
    Top8 t;
    t.i[0] = i0.x; t.i[1] = i0.y; t.i[2] = i0.z; t.i[3] = i0.w; t.i[4] = i1.x; t.i[5] = i1.y; t.i[6] = i1.z; t.i[7] = i1.w;
#pragma unroll
    for (int k = 0; k < DFU_KNN; ++k) t.d[k] = 0.f;  // the blend only reads the ids
    const float w[DFU_KNN] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    const DQ b = blend(a.blend_mode, t, w, a.real, a.dual);
    const V3 r = dq_transform_vertex(b, V3{qx, qy, qz});
    a.v_out[3 * (size_t) q] = r.x; a.v_out[3 * (size_t) q + 1] = r.y; a.v_out[3 * (size_t) q + 2] = r.z;
    if (a.n_in && a.n_out) {
        const V3 rn = a.normal_mode == DFU_NORMAL_REF ? dq_transform_vertex(b, nn) : dq_rotate(b, nn);
        a.n_out[3 * (size_t) q] = rn.x; a.n_out[3 * (size_t) q + 1] = rn.y; a.n_out[3 * (size_t) q + 2] = rn.z;
    }
}

int build_node_grid(dfu_warpfield* wf, cudaStream_t st) {
    NodeGrid& ng = wf->grid;
    if (ng.valid && ng.node_epoch == wf->node_epoch) return DFU_OK;
    if (!ng.desc) {
        DFU_CUDA_OK(cudaMalloc(&ng.desc, sizeof(GridDesc)));
        DFU_CUDA_OK(cudaMalloc(&ng.cell_start, ((size_t) DFU_GRID_MAX_CELLS + 1) * sizeof(int)));
        DFU_CUDA_OK(cudaMalloc(&ng.cell_count, (size_t) DFU_GRID_MAX_CELLS * sizeof(int)));
    }
    if (wf->N > ng.capacity) {
        if (ng.sorted) cudaFree(ng.sorted);
        ng.sorted = nullptr;
        ng.capacity = 0;
        DFU_CUDA_OK(cudaMalloc(&ng.sorted, (size_t) wf->N * sizeof(float4)));
        if (ng.occ) cudaFree(ng.occ);
        ng.occ = nullptr;
        DFU_CUDA_OK(cudaMalloc(&ng.occ, (size_t) wf->N * sizeof(int)));
        ng.capacity = wf->N;
    }
    grid_desc_kernel<<<1, 1024, 0, st>>>(wf->pos_w, wf->N, ng.desc);
    DFU_LAUNCH_OK();
    DFU_CUDA_OK(cudaMemsetAsync(ng.cell_count, 0, (size_t) DFU_GRID_MAX_CELLS * sizeof(int), st));
    grid_count_kernel<<<div_up(wf->N, 256), 256, 0, st>>>(wf->pos_w, wf->N, ng.desc, ng.cell_count);
    DFU_LAUNCH_OK();
    grid_scan_kernel<<<1, 1024, 0, st>>>(ng.cell_count, ng.desc, ng.cell_start);
    DFU_LAUNCH_OK();
    grid_fill_kernel<<<div_up(wf->N, 256), 256, 0, st>>>(wf->pos_w, wf->N, ng.desc, ng.cell_start, ng.cell_count, ng.sorted);
    DFU_LAUNCH_OK();
    grid_occ_kernel<<<div_up(DFU_GRID_MAX_CELLS, 256), 256, 0, st>>>(ng.cell_start, ng.desc, ng.occ);
    DFU_LAUNCH_OK();
    ng.node_epoch = wf->node_epoch;
    ng.valid = true;
    return DFU_OK;
}

template <int OP>
int launch_points(const dfu_warpfield* wf, PointArgs& a, cudaStream_t st) {
    a.pos_w = wf->pos_w;
    a.real = wf->real;
    a.dual = wf->dual;
    a.Npad = wf->Npad;
    if (a.Q == 0) return DFU_OK;
    // point queries go through the node grid (exact, tens of candidates per query); the smem-tiled brute force
    // stays for the brick-centre queries (most are far from every node), tiny node sets and DFU_POINT_KNN=brute
    if (OP != OP_BOUNDS && wf->N >= 64 && wf->grid.valid && wf->grid.node_epoch == wf->node_epoch &&
        ((reinterpret_cast<uintptr_t>(a.idx) | reinterpret_cast<uintptr_t>(a.dist2) | reinterpret_cast<uintptr_t>(a.wts) |
          reinterpret_cast<uintptr_t>(a.dq_out)) & 15) == 0) {
        const char* env = OP == OP_GRAPH ? getenv("DFU_GRAPH_LANES") : nullptr;
        const int lanes = OP == OP_GRAPH ? (env ? atoi(env) : 2) : 1;
        if (lanes == 4)
            points_grid_graph_kernel<4><<<div_up((long) a.Q * 4, 128), 128, 0, st>>>(a, wf->grid.desc, wf->grid.cell_start, wf->grid.sorted, wf->grid.occ);
        else if (lanes == 2)
            points_grid_graph_kernel<2><<<div_up((long) a.Q * 2, 128), 128, 0, st>>>(a, wf->grid.desc, wf->grid.cell_start, wf->grid.sorted, wf->grid.occ);
        else
            points_grid_kernel<OP><<<div_up(a.Q, 128), 128, 0, st>>>(a, wf->grid.desc, wf->grid.cell_start, wf->grid.sorted, wf->grid.occ);
        DFU_LAUNCH_OK();
        return DFU_OK;
    }
    static bool attr_set = false;  // per template instantiation
    if (!attr_set) {
        DFU_CUDA_OK(cudaFuncSetAttribute(points_kernel<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int) sizeof(KnnSmem)));
        attr_set = true;
    }
    points_kernel<OP><<<div_up(a.Q, QPB), 256, sizeof(KnnSmem), st>>>(a);
    DFU_LAUNCH_OK();
    return DFU_OK;
}

int ensure_capacity(dfu_warpfield* wf, int N) {
    const int Npad = (N + 31) / 32 * 32;
    if (Npad > wf->capacity) {
        if (wf->pos_w) cudaFree(wf->pos_w);
        if (wf->real) cudaFree(wf->real);
        if (wf->dual) cudaFree(wf->dual);
        wf->pos_w = wf->real = wf->dual = nullptr;
        wf->capacity = 0;
        DFU_CUDA_OK(cudaMalloc(&wf->pos_w, (size_t) Npad * sizeof(float4)));
        DFU_CUDA_OK(cudaMalloc(&wf->real, (size_t) Npad * sizeof(float4)));
        DFU_CUDA_OK(cudaMalloc(&wf->dual, (size_t) Npad * sizeof(float4)));
        wf->capacity = Npad;
    }
    if (!wf->flags) DFU_CUDA_OK(cudaMalloc(&wf->flags, 4 * sizeof(int)));
    wf->N = N;
    wf->Npad = Npad;
    return DFU_OK;
}

int ensure_staging(dfu_warpfield* wf, size_t floats) {
    if (floats > wf->staging_cap) {
        if (wf->staging) cudaFree(wf->staging);
        wf->staging = nullptr;
        wf->staging_cap = 0;
        DFU_CUDA_OK(cudaMalloc(&wf->staging, floats * sizeof(float)));
        wf->staging_cap = floats;
    }
    return DFU_OK;
}

}  // namespace

int dfu_wf_refresh_flags(dfu_warpfield* wf, cudaStream_t st) {
    node_flags_kernel<<<1, 1024, 0, st>>>(wf->pos_w, wf->real, wf->dual, wf->N, wf->flags);
    DFU_LAUNCH_OK();
    return DFU_OK;
}

// Acceleration data for the warped integrator: for every 8x8x8 brick of the volume, the squared distance
// from its centre to the 8th and to the 1st nearest node.  Depends only on node POSITIONS and the volume
// geometry, so it is cached until either changes (the reference rebuilds its KD-tree at the same moments,
// src/dynfu/warp_field.cpp:24-27,85-94).
int dfu_wf_build_brick_table(dfu_warpfield* wf, const int dims[3], const float voxel[3], int z0, int z1, cudaStream_t st) {
    BrickTable& bt = wf->bricks;
    const int zb0 = z0 / 8, zb1 = (z1 + 7) / 8;
    const bool geom = bt.valid && bt.dims[0] == dims[0] && bt.dims[1] == dims[1] && bt.dims[2] == dims[2] &&
                      bt.voxel[0] == voxel[0] && bt.voxel[1] == voxel[1] && bt.voxel[2] == voxel[2];
    const bool same = geom && bt.node_epoch == wf->node_epoch;
    const bool pool_covers = bt.pool_bricks == 0 || (zb0 >= bt.pool_zb0 && zb1 <= bt.pool_zb1);
    // the built[] flags survive a change of the node set when Warpfield::update has already cleared the bricks its
    // new nodes can reach (cache_epoch is then the current epoch although the bounds are stale)
    const bool cache_ok = geom && bt.cache_epoch == wf->node_epoch && pool_covers;
    if (same && cache_ok) return DFU_OK;
    const int bd[3] = {dims[0] / 8, dims[1] / 8, (dims[2] + 7) / 8};
    const size_t nb = (size_t) bd[0] * bd[1] * bd[2];
    if (!same) {
        if (nb > bt.capacity) {
            if (bt.bounds) cudaFree(bt.bounds);
            bt.bounds = nullptr;
            bt.capacity = 0;
            DFU_CUDA_OK(cudaMalloc(&bt.bounds, nb * sizeof(float2)));
            bt.capacity = nb;
        }
        PointArgs a{};
        a.Q = (int) nb;
        a.bounds = bt.bounds;
        for (int i = 0; i < 3; ++i) {
            a.bdim[i] = bd[i];
            a.voxel[i] = voxel[i];
            bt.dims[i] = dims[i];
            bt.voxel[i] = voxel[i];
        }
        int rc = launch_points<OP_BOUNDS>(wf, a, st);
        if (rc != DFU_OK) return rc;
    }
    // voxel cache for the brick planes of this z-slab: 24 KB per brick (512^3 full volume: 6 GiB of the 180 GB).
    // Budget DFU_VOXEL_CACHE_GB (default 16); DFU_VOXEL_KNN_CACHE=0 turns the cache off (every frame then recomputes
    // the per-voxel 8-NN and weights).
    const char* env = getenv("DFU_VOXEL_KNN_CACHE");
    const char* gb = getenv("DFU_VOXEL_CACHE_GB");
    const double budget = (gb ? atof(gb) : 16.0) * 1073741824.0;
    const size_t nbp = (size_t) bd[0] * bd[1] * (size_t) (zb1 - zb0);
    const bool want = !(env && env[0] == '0') && wf->N <= 65535 && (double) nbp * 24576.0 <= budget;
    if (!want || nbp > bt.pool_bricks) {
        if (bt.knn_pool) cudaFree(bt.knn_pool);
        if (bt.w_pool) cudaFree(bt.w_pool);
        if (bt.built) cudaFree(bt.built);
        bt.knn_pool = nullptr;
        bt.w_pool = nullptr;
        bt.built = nullptr;
        bt.pool_bricks = 0;
    }
    if (want) {
        const bool fresh = bt.pool_bricks == 0;
        if (fresh) {
            DFU_CUDA_OK(cudaMalloc(&bt.knn_pool, nbp * 8192ull));
            DFU_CUDA_OK(cudaMalloc(&bt.w_pool, nbp * 16384ull));
            DFU_CUDA_OK(cudaMalloc(&bt.built, nbp));
            bt.pool_bricks = nbp;
        }
        if (fresh || !cache_ok) {
            DFU_CUDA_OK(cudaMemsetAsync(bt.built, 0, bt.pool_bricks, st));
            bt.pool_zb0 = zb0;
            bt.pool_zb1 = zb1;
        }
    }
    bt.cache_epoch = wf->node_epoch;
    bt.node_epoch = wf->node_epoch;
    bt.valid = true;
    return DFU_OK;
}

// data graph + weights + (live - canon) for the solver; requires N >= 8
int dfu_wf_build_data_graph(const dfu_warpfield* wf, const float* canon, const float* live, int P, int32_t* nbr,
                            float* wts, float* dvec, int* deg, int32_t* rank, cudaStream_t st) {
    PointArgs a{};
    a.q = canon;
    a.Q = P;
    a.idx = nbr;
    a.wts = wts;
    a.live = live;
    a.dvec = dvec;
    a.deg = deg;
    a.rank = rank;
    return launch_points<OP_GRAPH>(wf, a, st);
}
// regularisation graph: 8-NN of every node among the nodes (includes itself; opt_solver.cpp:74-105)
int dfu_wf_build_node_graph(const dfu_warpfield* wf, int32_t* nnbr, cudaStream_t st) {
    PointArgs a{};
    a.Q = wf->N;
    a.idx = nnbr;
    return launch_points<OP_NODEGRAPH>(wf, a, st);
}

// ================================================================================================
extern "C" {

int dfu_warpfield_create(dfu_warpfield** out, int device) {
    DFU_REQUIRE(out != nullptr, DFU_ERR_INVALID, "out is NULL");
    int rc = dfu_device_check(device);
    if (rc != DFU_OK) return rc;
    *out = new dfu_warpfield();
    (*out)->device = device;
    return DFU_OK;
}

int dfu_warpfield_destroy(dfu_warpfield* wf) {
    if (!wf) return DFU_OK;
    DeviceGuard g(wf->device);
    cudaFree(wf->pos_w);
    cudaFree(wf->real);
    cudaFree(wf->dual);
    cudaFree(wf->flags);
    cudaFree(wf->staging);
    cudaFree(wf->bricks.bounds);
    cudaFree(wf->bricks.knn_pool);
    cudaFree(wf->bricks.w_pool);
    cudaFree(wf->bricks.built);
    cudaFree(wf->grid.desc);
    cudaFree(wf->grid.cell_start);
    cudaFree(wf->grid.cell_count);
    cudaFree(wf->grid.sorted);
    cudaFree(wf->grid.occ);
    delete wf;
    return DFU_OK;
}

int dfu_warpfield_init(dfu_warpfield* wf, float epsilon, const float* pos_xyz, const float* dq, const float* dg_w,
                       int N, dfu_stream stream) {
    DFU_REQUIRE(wf && pos_xyz && dq && dg_w, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(N >= 1, DFU_ERR_INVALID, "N must be >= 1");
    DFU_GUARD(wf->device);
    int rc = ensure_capacity(wf, N);
    if (rc != DFU_OK) return rc;
    wf->epsilon = epsilon;
    cudaStream_t st = as_stream(stream);
    pack_nodes_kernel<<<div_up(wf->Npad, 256), 256, 0, st>>>(pos_xyz, dq, dg_w, N, wf->Npad, wf->pos_w, wf->real,
                                                             wf->dual);
    DFU_LAUNCH_OK();
    wf->node_epoch++;
    wf->initialised = true;
    const char* env = getenv("DFU_POINT_KNN");
    wf->grid.valid = false;
    if (!(env && env[0] == 'b') && N >= 64) {
        rc = build_node_grid(wf, st);
        if (rc != DFU_OK) return rc;
    }
    return dfu_wf_refresh_flags(wf, st);
}

int dfu_warpfield_init_host(dfu_warpfield* wf, float epsilon, const float* pos_xyz_host, const float* dq_host,
                            const float* dg_w_host, int N, dfu_stream stream) {
    DFU_REQUIRE(wf && pos_xyz_host && dq_host && dg_w_host, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(N >= 1, DFU_ERR_INVALID, "N must be >= 1");
    DFU_GUARD(wf->device);
    int rc = ensure_staging(wf, (size_t) N * 12);
    if (rc != DFU_OK) return rc;
    cudaStream_t st = as_stream(stream);
    float* s = wf->staging;
    DFU_CUDA_OK(cudaMemcpyAsync(s, pos_xyz_host, (size_t) N * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    DFU_CUDA_OK(cudaMemcpyAsync(s + (size_t) N * 3, dq_host, (size_t) N * 8 * sizeof(float), cudaMemcpyHostToDevice, st));
    DFU_CUDA_OK(cudaMemcpyAsync(s + (size_t) N * 11, dg_w_host, (size_t) N * sizeof(float), cudaMemcpyHostToDevice, st));
    rc = dfu_warpfield_init(wf, epsilon, s, s + (size_t) N * 3, s + (size_t) N * 11, N, stream);
    if (rc != DFU_OK) return rc;
    DFU_CUDA_OK(cudaStreamSynchronize(st));  // host buffers may be reused on return
    return DFU_OK;
}

int dfu_warpfield_num_nodes(const dfu_warpfield* wf, int* N_host) {
    DFU_REQUIRE(wf && N_host, DFU_ERR_INVALID, "NULL argument");
    *N_host = wf->N;
    return DFU_OK;
}

int dfu_warpfield_get_nodes(const dfu_warpfield* wf, float* pos_xyz, float* dq, float* dg_w, dfu_stream stream) {
    DFU_REQUIRE(wf, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(wf->initialised, DFU_ERR_NOT_INIT, "warp field not initialised");
    DFU_GUARD(wf->device);
    unpack_nodes_kernel<<<div_up(wf->N, 256), 256, 0, as_stream(stream)>>>(wf->pos_w, wf->real, wf->dual, wf->N,
                                                                           pos_xyz, dq, dg_w);
    DFU_LAUNCH_OK();
    return DFU_OK;
}

int dfu_warpfield_get_nodes_host(const dfu_warpfield* wf, float* pos_xyz_host, float* dq_host, float* dg_w_host,
                                 dfu_stream stream) {
    DFU_REQUIRE(wf, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(wf->initialised, DFU_ERR_NOT_INIT, "warp field not initialised");
    DFU_GUARD(wf->device);
    dfu_warpfield* w = const_cast<dfu_warpfield*>(wf);
    const size_t N = (size_t) wf->N;
    int rc = ensure_staging(w, N * 12);
    if (rc != DFU_OK) return rc;
    float* s = w->staging;
    rc = dfu_warpfield_get_nodes(wf, s, s + N * 3, s + N * 11, stream);
    if (rc != DFU_OK) return rc;
    cudaStream_t st = as_stream(stream);
    if (pos_xyz_host) DFU_CUDA_OK(cudaMemcpyAsync(pos_xyz_host, s, N * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (dq_host) DFU_CUDA_OK(cudaMemcpyAsync(dq_host, s + N * 3, N * 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (dg_w_host) DFU_CUDA_OK(cudaMemcpyAsync(dg_w_host, s + N * 11, N * sizeof(float), cudaMemcpyDeviceToHost, st));
    DFU_CUDA_OK(cudaStreamSynchronize(st));
    return DFU_OK;
}

int dfu_warpfield_set_transforms(dfu_warpfield* wf, const float* dq, dfu_stream stream) {
    DFU_REQUIRE(wf && dq, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(wf->initialised, DFU_ERR_NOT_INIT, "warp field not initialised");
    DFU_GUARD(wf->device);
    cudaStream_t st = as_stream(stream);
    pack_nodes_kernel<<<div_up(wf->Npad, 256), 256, 0, st>>>(nullptr, dq, nullptr, wf->N, wf->Npad, wf->pos_w,
                                                             wf->real, wf->dual);
    DFU_LAUNCH_OK();
    return dfu_wf_refresh_flags(wf, st);
}

int dfu_warpfield_set_transforms_host(dfu_warpfield* wf, const float* dq_host, dfu_stream stream) {
    DFU_REQUIRE(wf && dq_host, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(wf->initialised, DFU_ERR_NOT_INIT, "warp field not initialised");
    DFU_GUARD(wf->device);
    int rc = ensure_staging(wf, (size_t) wf->N * 12);
    if (rc != DFU_OK) return rc;
    cudaStream_t st = as_stream(stream);
    DFU_CUDA_OK(cudaMemcpyAsync(wf->staging, dq_host, (size_t) wf->N * 8 * sizeof(float), cudaMemcpyHostToDevice, st));
    rc = dfu_warpfield_set_transforms(wf, wf->staging, stream);
    if (rc != DFU_OK) return rc;
    DFU_CUDA_OK(cudaStreamSynchronize(st));
    return DFU_OK;
}

int dfu_warpfield_update_translations(dfu_warpfield* wf, const float* t_xyz, dfu_stream stream) {
    DFU_REQUIRE(wf && t_xyz, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(wf->initialised, DFU_ERR_NOT_INIT, "warp field not initialised");
    DFU_GUARD(wf->device);
    cudaStream_t st = as_stream(stream);
    update_translations_kernel<<<div_up(wf->N, 256), 256, 0, st>>>(wf->real, wf->dual, t_xyz, wf->N);
    DFU_LAUNCH_OK();
    return dfu_wf_refresh_flags(wf, st);
}

int dfu_warpfield_knn(const dfu_warpfield* wf, const float* q_xyz, int Q, int32_t* idx, float* dist2,
                      dfu_stream stream) {
    DFU_REQUIRE(wf && (Q == 0 || (q_xyz && idx)), DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(Q >= 0, DFU_ERR_INVALID, "negative Q");
    DFU_REQUIRE(wf->initialised, DFU_ERR_NOT_INIT, "warp field not initialised");
    DFU_GUARD(wf->device);
    PointArgs a{};
    a.q = q_xyz;
    a.Q = Q;
    a.idx = idx;
    a.dist2 = dist2;
    return launch_points<OP_KNN>(wf, a, as_stream(stream));
}

int dfu_warpfield_blend(const dfu_warpfield* wf, const float* p_xyz, int Q, float* dq_out, int blend_mode,
                        dfu_stream stream) {
    DFU_REQUIRE(wf && (Q == 0 || (p_xyz && dq_out)), DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(Q >= 0, DFU_ERR_INVALID, "negative Q");
    DFU_REQUIRE(blend_mode == DFU_BLEND_REF_COMPOSE || blend_mode == DFU_BLEND_DQB_SUM, DFU_ERR_INVALID, "bad blend_mode");
    DFU_REQUIRE(wf->initialised, DFU_ERR_NOT_INIT, "warp field not initialised");
    DFU_GUARD(wf->device);
    PointArgs a{};
    a.q = p_xyz;
    a.Q = Q;
    a.dq_out = dq_out;
    a.blend_mode = blend_mode;
    return launch_points<OP_BLEND>(wf, a, as_stream(stream));
}

int dfu_warpfield_warp(const dfu_warpfield* wf, const float* v_xyz, const float* n_xyz, int P, float* v_out,
                       float* n_out, int blend_mode, int normal_mode, dfu_stream stream) {
    DFU_REQUIRE(wf && (P == 0 || (v_xyz && v_out)), DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(P >= 0, DFU_ERR_INVALID, "negative P");
    DFU_REQUIRE(blend_mode == DFU_BLEND_REF_COMPOSE || blend_mode == DFU_BLEND_DQB_SUM, DFU_ERR_INVALID, "bad blend_mode");
    DFU_REQUIRE(normal_mode == DFU_NORMAL_REF || normal_mode == DFU_NORMAL_ROTATE_ONLY, DFU_ERR_INVALID, "bad normal_mode");
    DFU_REQUIRE(wf->initialised, DFU_ERR_NOT_INIT, "warp field not initialised");
    DFU_GUARD(wf->device);
    PointArgs a{};
    a.q = v_xyz;
    a.Q = P;
    a.n_in = n_xyz;
    a.v_out = v_out;
    a.n_out = n_out;
    a.blend_mode = blend_mode;
    a.normal_mode = normal_mode;
    return launch_points<OP_WARP>(wf, a, as_stream(stream));
}

struct dfu_pointcache {
    int device = 0;
    int32_t* ids = nullptr;
    float* wts = nullptr;
    size_t cap = 0;
    // what the cache was filled for
    const dfu_warpfield* wf = nullptr;
    uint64_t node_epoch = 0;
    unsigned long long version = 0;
    const float* v = nullptr;
    int P = -1;
};

int dfu_pointcache_create(dfu_pointcache** out, int device) {
    DFU_REQUIRE(out != nullptr, DFU_ERR_INVALID, "out is NULL");
    int rc = dfu_device_check(device);
    if (rc != DFU_OK) return rc;
    *out = new dfu_pointcache();
    (*out)->device = device;
    return DFU_OK;
}

int dfu_pointcache_destroy(dfu_pointcache* c) {
    if (!c) return DFU_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    cudaFree(c->ids);
    cudaFree(c->wts);
    cudaSetDevice(prev);
    delete c;
    return DFU_OK;
}

int dfu_warpfield_warp_cached(const dfu_warpfield* wf, dfu_pointcache* cache, unsigned long long points_version, const float* v_xyz,
                              const float* n_xyz, int P, float* v_out, float* n_out, int blend_mode, int normal_mode,
                              dfu_stream stream) {
    DFU_REQUIRE(wf && cache && (P == 0 || (v_xyz && v_out)), DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(P >= 0, DFU_ERR_INVALID, "negative P");
    DFU_REQUIRE(blend_mode == DFU_BLEND_REF_COMPOSE || blend_mode == DFU_BLEND_DQB_SUM, DFU_ERR_INVALID, "bad blend_mode");
    DFU_REQUIRE(normal_mode == DFU_NORMAL_REF || normal_mode == DFU_NORMAL_ROTATE_ONLY, DFU_ERR_INVALID, "bad normal_mode");
    DFU_REQUIRE(wf->initialised, DFU_ERR_NOT_INIT, "warp field not initialised");
    DFU_REQUIRE(v_out != v_xyz, DFU_ERR_INVALID, "the cached warp cannot run in place: the cache is keyed on the input points");
    if (P == 0) return DFU_OK;
    DFU_GUARD(wf->device);
    cudaStream_t st = as_stream(stream);
    const bool hit = cache->wf == wf && cache->node_epoch == wf->node_epoch && cache->version == points_version &&
                     cache->v == v_xyz && cache->P == P;
    if (!hit) {
        if ((size_t) P > cache->cap) {
            cudaFree(cache->ids);
            cudaFree(cache->wts);
            cache->ids = nullptr;
            cache->wts = nullptr;
            cache->cap = 0;
            DFU_CUDA_OK(cudaMalloc(&cache->ids, (size_t) P * DFU_KNN * sizeof(int32_t)));
            DFU_CUDA_OK(cudaMalloc(&cache->wts, (size_t) P * DFU_KNN * sizeof(float)));
            cache->cap = (size_t) P;
        }
        cache->P = -1;
        PointArgs g{};  // neighbours + weights of every point: the solver's graph kernel without its extras
        g.q = v_xyz;
        g.Q = P;
        g.idx = cache->ids;
        g.wts = cache->wts;
        int rc = launch_points<OP_GRAPH>(wf, g, st);
        if (rc != DFU_OK) return rc;
        cache->wf = wf;
        cache->node_epoch = wf->node_epoch;
        cache->version = points_version;
        cache->v = v_xyz;
        cache->P = P;
    }
    PointArgs a{};
    a.q = v_xyz;
    a.Q = P;
    a.n_in = n_xyz;
    a.v_out = v_out;
    a.n_out = n_out;
    a.blend_mode = blend_mode;
    a.normal_mode = normal_mode;
    a.real = wf->real;
    a.dual = wf->dual;
    warp_cached_kernel<<<div_up(P, 256), 256, 0, st>>>(a, cache->ids, cache->wts);
    DFU_LAUNCH_OK();
    return DFU_OK;
}

}  // extern "C"
