// The callers and data formats on the input side of the hot path:
//   * cuda::computePointNormals  (src/kfusion/imgproc.cpp:27-36 -> src/kfusion/cuda/imgproc.cu:187-226):
//     depth image -> per-pixel camera-space points and normals (float4 images, NaN where invalid);
//   * compaction of the valid pixels into packed vertex / normal arrays (what the reference obtains by downloading
//     the marching-cubes triangles, src/dynfu/dyn_fusion.cpp:120-134), optionally moved into another frame;
//   * DynFusion::findCorrespondingFrame (src/dynfu/dyn_fusion.cpp:212-242): for every live vertex the nearest
//     (warped) canonical vertex, replacing the second nanoflann KD-tree the reference builds every frame by a
//     uniform grid over the canonical vertices searched in shells (exact, key (dist2, index)).
#include <math_constants.h>

#include <algorithm>

#include "dfu_internal.h"
#include "dfu_math.cuh"
#include "scan.cuh"

using namespace dfu;

namespace {

// ---- points + normals ----------------------------------------------------------------------------------------
// Reprojector::operator() (include/kfusion/cuda/device.hpp:50-54): x = z * (u - cx) * finv.x, left to right
DFU_DEV V3 reproject(int u, int v, float z, float finvx, float finvy, float cx, float cy) {
    return V3{fmul(fmul(z, fsub((float) u, cx)), finvx), fmul(fmul(z, fsub((float) v, cy)), finvy), z};
}

// points_normals_kernel (src/kfusion/cuda/imgproc.cu:187-215).  The reference normalises with the approximate
// rsqrt(); here n = cross / sqrt(dot) in IEEE arithmetic so that the CPU oracle can reproduce it bit for bit.
__global__ void points_normals_kernel(const uint16_t* __restrict__ depth, size_t dpitch, int rows, int cols, float finvx, float finvy,
                                      float cx, float cy, float4* __restrict__ points, size_t ppitch, float4* __restrict__ normals,
                                      size_t npitch) {
    const int x = threadIdx.x + blockIdx.x * blockDim.x;
    const int y = threadIdx.y + blockIdx.y * blockDim.y;
    if (x >= cols || y >= rows) return;
    const float qnan = CUDART_NAN_F;
    float4 P = make_float4(qnan, qnan, qnan, qnan), Nn = P;
    if (x < cols - 1 && y < rows - 1) {
        const uint16_t* r0 = reinterpret_cast<const uint16_t*>(reinterpret_cast<const char*>(depth) + (size_t) y * dpitch);
        const uint16_t* r1 = reinterpret_cast<const uint16_t*>(reinterpret_cast<const char*>(depth) + (size_t) (y + 1) * dpitch);
        const float z00 = fmul((float) r0[x], 0.001f), z01 = fmul((float) r0[x + 1], 0.001f), z10 = fmul((float) r1[x], 0.001f);
        if (fmul(fmul(z00, z01), z10) != 0.f) {
            const V3 v00 = reproject(x, y, z00, finvx, finvy, cx, cy);
            const V3 v01 = reproject(x + 1, y, z01, finvx, finvy, cx, cy);
            const V3 v10 = reproject(x, y + 1, z10, finvx, finvy, cx, cy);
            const V3 c = cross(vsub(v01, v00), vsub(v10, v00));
            const float len = __fsqrt_rn(fadd(fadd(fmul(c.x, c.x), fmul(c.y, c.y)), fmul(c.z, c.z)));
            Nn = make_float4(-__fdiv_rn(c.x, len), -__fdiv_rn(c.y, len), -__fdiv_rn(c.z, len), 0.f);
            P = make_float4(v00.x, v00.y, v00.z, 0.f);
        }
    }
    *reinterpret_cast<float4*>(reinterpret_cast<char*>(points) + (size_t) y * ppitch + sizeof(float4) * x) = P;
    *reinterpret_cast<float4*>(reinterpret_cast<char*>(normals) + (size_t) y * npitch + sizeof(float4) * x) = Nn;
}

// ---- deterministic compaction of the valid pixels, raster order --------------------------------------------------
struct Xform {
    float m[12];
    int on;
};
DFU_DEV bool valid_px(const float4* img, size_t pitch, int x, int y) {
    const float4 p = *reinterpret_cast<const float4*>(reinterpret_cast<const char*>(img) + (size_t) y * pitch + sizeof(float4) * x);
    return p.x == p.x && p.y == p.y && p.z == p.z;  // not NaN
}
// one warp per 32-pixel chunk of a row: number of valid pixels
__global__ void chunk_count_kernel(const float4* __restrict__ points, size_t ppitch, const float4* __restrict__ normals, size_t npitch,
                                   int rows, int cols, int cpr, int* __restrict__ counts) {
    const int chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (chunk >= rows * cpr) return;
    const int y = chunk / cpr, x = (chunk - y * cpr) * 32 + lane;
    const bool ok = x < cols && valid_px(points, ppitch, x, y) && (!normals || valid_px(normals, npitch, x, y));
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) counts[chunk] = __popc(m);
}
// one warp per chunk writes its valid pixels at the chunk's offset, in x order; optional rigid transform of the points
// (rotation only for the normals)
__global__ void chunk_emit_kernel(const float4* __restrict__ points, size_t ppitch, const float4* __restrict__ normals, size_t npitch,
                                  int rows, int cols, int cpr, const int* __restrict__ offsets, int capacity, const Xform xf,
                                  float* __restrict__ out_v, float* __restrict__ out_n) {
    const int chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (chunk >= rows * cpr) return;
    const int y = chunk / cpr, x = (chunk - y * cpr) * 32 + lane;
    const bool ok = x < cols && valid_px(points, ppitch, x, y) && (!normals || valid_px(normals, npitch, x, y));
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    const int slot = offsets[chunk] + __popc(m & ((1u << lane) - 1u));
    if (!ok || slot >= capacity) return;
    float4 p = *reinterpret_cast<const float4*>(reinterpret_cast<const char*>(points) + (size_t) y * ppitch + sizeof(float4) * x);
    float4 n = make_float4(0.f, 0.f, 0.f, 0.f);
    if (normals) n = *reinterpret_cast<const float4*>(reinterpret_cast<const char*>(normals) + (size_t) y * npitch + sizeof(float4) * x);
    if (xf.on) {  // p' = R p + t with the fmaf chain of the integrator; n' = R n
        const float* a = xf.m;
        const float px = __fmaf_rn(a[2], p.z, __fmaf_rn(a[1], p.y, __fmaf_rn(a[0], p.x, a[9])));
        const float py = __fmaf_rn(a[5], p.z, __fmaf_rn(a[4], p.y, __fmaf_rn(a[3], p.x, a[10])));
        const float pz = __fmaf_rn(a[8], p.z, __fmaf_rn(a[7], p.y, __fmaf_rn(a[6], p.x, a[11])));
        const float nx = __fmaf_rn(a[2], n.z, __fmaf_rn(a[1], n.y, fmul(a[0], n.x)));
        const float ny = __fmaf_rn(a[5], n.z, __fmaf_rn(a[4], n.y, fmul(a[3], n.x)));
        const float nz = __fmaf_rn(a[8], n.z, __fmaf_rn(a[7], n.y, fmul(a[6], n.x)));
        p = make_float4(px, py, pz, 0.f);
        n = make_float4(nx, ny, nz, 0.f);
    }
    out_v[3 * (size_t) slot] = p.x; out_v[3 * (size_t) slot + 1] = p.y; out_v[3 * (size_t) slot + 2] = p.z;
    if (out_n) {
        out_n[3 * (size_t) slot] = n.x; out_n[3 * (size_t) slot + 1] = n.y; out_n[3 * (size_t) slot + 2] = n.z;
    }
}

// ---- point index: uniform grid over an arbitrary point set, exact 1-NN ---------------------------------------
struct PGrid {
    float ox, oy, oz, h, inv_h;
    int nx, ny, nz;
    int n_occ;   // non-empty cells of the current binning
    int n_pts;
};
constexpr int PG_MAX_DIM = 160;
constexpr int PG_MAX_CELLS = PG_MAX_DIM * PG_MAX_DIM * PG_MAX_DIM;

// bbox[0..2] = min (ordered uint), bbox[3..5] = max, bbox[6] = CTAs done; the last CTA derives the first-guess grid
__global__ void __launch_bounds__(256) pg_bbox_kernel(const float* __restrict__ pts, int P, unsigned* __restrict__ bbox,
                                                      PGrid* __restrict__ g) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x)
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
            atomicMin(&bbox[c], f2ord(mn[c]));
            atomicMax(&bbox[3 + c], f2ord(mx[c]));
        }
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&bbox[6], 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last || threadIdx.x != 0) return;
    __threadfence();
    for (int c = 0; c < 3; ++c) {
        mn[c] = ord2f(atomicOr(&bbox[c], 0u));
        mx[c] = ord2f(atomicOr(&bbox[3 + c], 0u));
    }
    const float ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
    const float ext = fmaxf(ex, fmaxf(ey, ez));
    // first guess: the points sample a surface about as large as the biggest face of their bounding box
    const float area = fmaxf(ex * ey, fmaxf(ey * ez, ex * ez));
    float h = 2.5f * sqrtf(fmaxf(area, 1e-12f) / (float) max(P, 1));
    h = fmaxf(h, ext / (float) (PG_MAX_DIM - 1));
    h = fmaxf(h, 1e-6f);
    g->ox = mn[0]; g->oy = mn[1]; g->oz = mn[2];
    g->h = h;
    g->inv_h = 1.f / h;
    g->nx = min(PG_MAX_DIM, (int) (ex * g->inv_h) + 1);
    g->ny = min(PG_MAX_DIM, (int) (ey * g->inv_h) + 1);
    g->nz = min(PG_MAX_DIM, (int) (ez * g->inv_h) + 1);
    g->n_occ = 0;
    g->n_pts = P;
}
DFU_DEV int pg_coord(float v, float o, float inv_h, int n) { return min(n - 1, max(0, (int) floorf((v - o) * inv_h))); }
DFU_DEV int pg_cell(const PGrid& g, float x, float y, float z) {
    return pg_coord(x, g.ox, g.inv_h, g.nx) + g.nx * (pg_coord(y, g.oy, g.inv_h, g.ny) + g.ny * pg_coord(z, g.oz, g.inv_h, g.nz));
}
// the grid size lives on the device: fixed launch shapes, CTAs past the end leave at once
__global__ void pg_zero_kernel(const PGrid* __restrict__ gd, int* __restrict__ count) {
    const int n = gd->nx * gd->ny * gd->nz;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) count[i] = 0;
}
__global__ void pg_count_kernel(const float* __restrict__ pts, int P, const PGrid* __restrict__ gd, int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const PGrid g = *gd;
    atomicAdd(&count[pg_cell(g, pts[3 * (size_t) i], pts[3 * (size_t) i + 1], pts[3 * (size_t) i + 2])], 1);
}
// pass 0: number of non-empty cells of the first-guess grid
__global__ void __launch_bounds__(256) pg_occupancy_kernel(PGrid* __restrict__ gd, const int* __restrict__ count) {
    const int n = gd->nx * gd->ny * gd->nz;
    int occ = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) occ += count[i] > 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) occ += __shfl_xor_sync(0xffffffffu, occ, o);
    if ((threadIdx.x & 31) == 0 && occ) atomicAdd(&gd->n_occ, occ);
}
// rescale the cell so that a non-empty cell holds ~6 points (the bounding-box guess can be far off)
__global__ void pg_adjust_kernel(PGrid* __restrict__ g) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float occ = (float) g->n_pts / (float) max(g->n_occ, 1);
    const float ex = (float) (g->nx) * g->h, ey = (float) (g->ny) * g->h, ez = (float) (g->nz) * g->h;
    float h = g->h * sqrtf(6.f / fmaxf(occ, 1e-3f));  // surface-like scaling: occupancy ~ h^2
    h = fmaxf(h, fmaxf(ex, fmaxf(ey, ez)) / (float) (PG_MAX_DIM - 1));
    h = fmaxf(h, 1e-6f);
    g->h = h;
    g->inv_h = 1.f / h;
    g->nx = min(PG_MAX_DIM, (int) (ex * g->inv_h) + 1);
    g->ny = min(PG_MAX_DIM, (int) (ey * g->inv_h) + 1);
    g->nz = min(PG_MAX_DIM, (int) (ez * g->inv_h) + 1);
    g->n_occ = 0;
}
// three-kernel exclusive scan of the cell counts: per-1024-cell sums, scan of the sums, per-cell offsets (+ the list
// of non-empty cells, in no particular order: the search key (dist2, idx) makes the result order independent)
constexpr int PG_SCAN_CHUNK = 1024;
constexpr int PG_MAX_CHUNKS = (PG_MAX_CELLS + PG_SCAN_CHUNK - 1) / PG_SCAN_CHUNK;
__global__ void __launch_bounds__(256) pg_chunk_sum_kernel(const PGrid* __restrict__ gd, const int* __restrict__ count,
                                                           int* __restrict__ chunk_sum) {
    const int n = gd->nx * gd->ny * gd->nz;
    const int base = blockIdx.x * PG_SCAN_CHUNK;
    if (base >= n) return;
    int s = 0;
#pragma unroll
    for (int k = 0; k < PG_SCAN_CHUNK / 256; ++k) {
        const int i = base + k * 256 + threadIdx.x;
        s += i < n ? count[i] : 0;
    }
    __shared__ int sh[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) chunk_sum[blockIdx.x] = sh[0] + sh[1] + sh[2] + sh[3] + sh[4] + sh[5] + sh[6] + sh[7];
}
__global__ void __launch_bounds__(1024) pg_chunk_scan_kernel(const PGrid* __restrict__ gd, int* __restrict__ chunk_sum) {
    __shared__ int sh[1024];
    const int n = (gd->nx * gd->ny * gd->nz + PG_SCAN_CHUNK - 1) / PG_SCAN_CHUNK;
    const int per = (n + 1023) / 1024;
    const int lo = min(n, (int) threadIdx.x * per), hi = min(n, lo + per);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += chunk_sum[i];
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
        const int c = chunk_sum[i];
        chunk_sum[i] = run;
        run += c;
    }
}
__global__ void __launch_bounds__(256) pg_offsets_kernel(PGrid* __restrict__ gd, const int* __restrict__ count,
                                                         const int* __restrict__ chunk_sum, int* __restrict__ start,
                                                         int* __restrict__ occ) {
    const int n = gd->nx * gd->ny * gd->nz;
    const int base = blockIdx.x * PG_SCAN_CHUNK;
    if (base >= n) return;
    // each thread owns 4 consecutive cells
    const int i0 = base + threadIdx.x * 4;
    int c[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c[k] = i0 + k < n ? count[i0 + k] : 0;
        s += c[k];
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    __shared__ int wsum[8];
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    int run = chunk_sum[blockIdx.x] + inc - s;
    for (int w = 0; w < wid; ++w) run += wsum[w];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (i0 + k < n) {
            start[i0 + k] = run;
            if (c[k] > 0) occ[atomicAdd(&gd->n_occ, 1)] = i0 + k;
        }
        run += c[k];
    }
    if (i0 <= n - 1 && n - 1 < i0 + 4) start[n] = run;
}
__global__ void pg_fill_kernel(const float* __restrict__ pts, int P, const PGrid* __restrict__ gd, const int* __restrict__ start,
                               int* __restrict__ count, float4* __restrict__ sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const PGrid g = *gd;
    const float x = pts[3 * (size_t) i], y = pts[3 * (size_t) i + 1], z = pts[3 * (size_t) i + 2];
    const int c = pg_cell(g, x, y, z);
    const int slot = atomicSub(&count[c], 1) - 1;  // order inside a cell is irrelevant: the search key is (dist2, idx)
    sorted[start[c] + slot] = make_float4(x, y, z, __int_as_float(i));
}

DFU_DEV void nn_visit(const int* __restrict__ start, const float4* __restrict__ sorted, int cell, float qx, float qy, float qz, float& bd,
                      int& bi) {
    const int lo = __ldg(&start[cell]), hi = __ldg(&start[cell + 1]);
    for (int j = lo; j < hi; ++j) {
        const float4 p = __ldg(&sorted[j]);
        const float d = dist2(qx, qy, qz, p.x, p.y, p.z);
        const int idx = __float_as_int(p.w);
        if (d < bd || (d == bd && idx < bi)) {
            bd = d;
            bi = idx;
        }
    }
}

// exact nearest neighbour (key (dist2, idx)) of every query: shells around the query's cell, then -- for queries far
// from the point set in units of the cell size -- a sweep over the non-empty cells pruned by box distance
__global__ void __launch_bounds__(128) pg_nearest_kernel(const float* __restrict__ q, int Q, const PGrid* __restrict__ gd,
                                                         const int* __restrict__ start, const float4* __restrict__ sorted,
                                                         const int* __restrict__ occ, int32_t* __restrict__ idx_out,
                                                         float* __restrict__ d2_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q) return;
    const PGrid g = *gd;
    const float qx = q[3 * (size_t) i], qy = q[3 * (size_t) i + 1], qz = q[3 * (size_t) i + 2];
    float bd = INFINITY;
    int bi = 0x7fffffff;
    const int cx = pg_coord(qx, g.ox, g.inv_h, g.nx), cy = pg_coord(qy, g.oy, g.inv_h, g.ny), cz = pg_coord(qz, g.oz, g.inv_h, g.nz);
    bool settled = false;
    int r = 0, rd = -1;
    for (; (2 * r + 1) * (2 * r + 1) * (2 * r + 1) <= 2 * g.n_occ + 27; ++r) {
        rd = r;
        const int z0 = max(0, cz - r), z1 = min(g.nz - 1, cz + r);
        const int y0 = max(0, cy - r), y1 = min(g.ny - 1, cy + r);
        const int x0 = max(0, cx - r), x1 = min(g.nx - 1, cx + r);
        for (int z = z0; z <= z1; ++z)
            for (int y = y0; y <= y1; ++y) {
                const bool face = (abs(z - cz) == r) || (abs(y - cy) == r);
                const int row = g.nx * (y + g.ny * z);
                if (face) {
                    for (int x = x0; x <= x1; ++x) nn_visit(start, sorted, row + x, qx, qy, qz, bd, bi);
                } else {
                    if (cx - r >= 0) nn_visit(start, sorted, row + cx - r, qx, qy, qz, bd, bi);
                    if (cx + r < g.nx) nn_visit(start, sorted, row + cx + r, qx, qy, qz, bd, bi);
                }
            }
        float L = INFINITY;
        if (cx - r > 0) L = fminf(L, qx - (g.ox + (float) (cx - r) * g.h));
        if (cx + r < g.nx - 1) L = fminf(L, (g.ox + (float) (cx + r + 1) * g.h) - qx);
        if (cy - r > 0) L = fminf(L, qy - (g.oy + (float) (cy - r) * g.h));
        if (cy + r < g.ny - 1) L = fminf(L, (g.oy + (float) (cy + r + 1) * g.h) - qy);
        if (cz - r > 0) L = fminf(L, qz - (g.oz + (float) (cz - r) * g.h));
        if (cz + r < g.nz - 1) L = fminf(L, (g.oz + (float) (cz + r + 1) * g.h) - qz);
        const float Ls = L - 1e-5f - 1e-5f * g.h;  // binning / face coordinates are rounded: margin >> ulp
        if (L == INFINITY || (Ls > 0.f && bd < Ls * Ls)) {
            settled = true;
            break;
        }
    }
    if (!settled) {
        for (int e = 0; e < g.n_occ; ++e) {
            const int c = __ldg(&occ[e]);
            const int ix = c % g.nx, iy = (c / g.nx) % g.ny, iz = c / (g.nx * g.ny);
            if (abs(ix - cx) <= rd && abs(iy - cy) <= rd && abs(iz - cz) <= rd) continue;
            const float lx = g.ox + (float) ix * g.h, ly = g.oy + (float) iy * g.h, lz = g.oz + (float) iz * g.h;
            const float ddx = fmaxf(0.f, fmaxf(lx - qx, qx - (lx + g.h))), ddy = fmaxf(0.f, fmaxf(ly - qy, qy - (ly + g.h))),
                        ddz = fmaxf(0.f, fmaxf(lz - qz, qz - (lz + g.h)));
            const float dl = fmaxf(0.f, sqrtf(ddx * ddx + ddy * ddy + ddz * ddz) - 1e-5f - 1e-5f * g.h);
            if (dl * dl <= bd) nn_visit(start, sorted, c, qx, qy, qz, bd, bi);
        }
    }
    idx_out[i] = bi == 0x7fffffff ? -1 : bi;
    if (d2_out) d2_out[i] = bd;
}

// correspondingCanonicalVertices.push_back(canonicalVertices[index]) (dyn_fusion.cpp:232-239)
__global__ void gather_rows_kernel(const float* __restrict__ src_v, const float* __restrict__ src_n, const int32_t* __restrict__ idx,
                                   int Q, float* __restrict__ out_v, float* __restrict__ out_n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q) return;
    const int j = idx[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        out_v[3 * (size_t) i + c] = j >= 0 ? src_v[3 * (size_t) j + c] : CUDART_NAN_F;
        if (src_n && out_n) out_n[3 * (size_t) i + c] = j >= 0 ? src_n[3 * (size_t) j + c] : CUDART_NAN_F;
    }
}

}  // namespace

struct dfu_pointindex {
    int device = 0;
    int P = 0, capacity = 0;
    PGrid* desc = nullptr;
    int *cell_start = nullptr, *cell_count = nullptr, *occ = nullptr, *cursor = nullptr;
    float4* sorted = nullptr;
    int32_t* tmp_idx = nullptr;
    int tmp_cap = 0;
    bool built = false;
};

extern "C" {

int dfu_compute_points_normals(const uint16_t* depth, size_t depth_pitch_bytes, int rows, int cols, const float intr_host[4],
                               float* points4, size_t points_pitch_bytes, float* normals4, size_t normals_pitch_bytes,
                               dfu_stream stream) {
    DFU_REQUIRE(depth && intr_host && points4 && normals4, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(rows > 0 && cols > 0, DFU_ERR_INVALID, "bad image size");
    (void) cudaGetLastError();
    dim3 block(32, 8), grid(div_up(cols, 32), div_up(rows, 8));
    points_normals_kernel<<<grid, block, 0, as_stream(stream)>>>(depth, depth_pitch_bytes, rows, cols, 1.f / intr_host[0],
                                                                 1.f / intr_host[1], intr_host[2], intr_host[3],
                                                                 reinterpret_cast<float4*>(points4), points_pitch_bytes,
                                                                 reinterpret_cast<float4*>(normals4), normals_pitch_bytes);
    DFU_LAUNCH_OK();
    return DFU_OK;
}

int dfu_compact_points(const float* points4, size_t points_pitch_bytes, const float* normals4, size_t normals_pitch_bytes, int rows,
                       int cols, const float xform_host[12], float* out_v, float* out_n, int capacity, int* count_out,
                       dfu_stream stream) {
    DFU_REQUIRE(points4 && out_v && count_out, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(rows > 0 && cols > 0 && capacity >= 0, DFU_ERR_INVALID, "bad size");
    (void) cudaGetLastError();
    cudaStream_t st = as_stream(stream);
    const float4* p4 = reinterpret_cast<const float4*>(points4);
    const float4* n4 = reinterpret_cast<const float4*>(normals4);
    Xform xf = {};
    if (xform_host) {
        for (int i = 0; i < 12; ++i) xf.m[i] = xform_host[i];
        xf.on = 1;
    }
    const int cpr = div_up(cols, 32), chunks = rows * cpr;
    int* offsets = nullptr;
    DFU_CUDA_OK(scratch_alloc((void**) &offsets, (size_t) chunks * sizeof(int), st));
    chunk_count_kernel<<<div_up(chunks, 8), 256, 0, st>>>(p4, points_pitch_bytes, n4, normals_pitch_bytes, rows, cols, cpr, offsets);
    DFU_LAUNCH_OK();
    small_scan_kernel<<<1, 1024, 0, st>>>(offsets, chunks, count_out);
    DFU_LAUNCH_OK();
    chunk_emit_kernel<<<div_up(chunks, 8), 256, 0, st>>>(p4, points_pitch_bytes, n4, normals_pitch_bytes, rows, cols, cpr, offsets,
                                                         capacity, xf, out_v, out_n);
    DFU_LAUNCH_OK();
    DFU_CUDA_OK(cudaFreeAsync(offsets, st));
    return DFU_OK;
}

int dfu_pointindex_create(dfu_pointindex** out, int device) {
    DFU_REQUIRE(out != nullptr, DFU_ERR_INVALID, "out is NULL");
    int rc = dfu_device_check(device);
    if (rc != DFU_OK) return rc;
    *out = new dfu_pointindex();
    (*out)->device = device;
    return DFU_OK;
}

int dfu_pointindex_destroy(dfu_pointindex* pi) {
    if (!pi) return DFU_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(pi->device);
    cudaFree(pi->desc); cudaFree(pi->cell_start); cudaFree(pi->cell_count); cudaFree(pi->occ); cudaFree(pi->cursor);
    cudaFree(pi->sorted); cudaFree(pi->tmp_idx);
    cudaSetDevice(prev);
    delete pi;
    return DFU_OK;
}

int dfu_pointindex_build(dfu_pointindex* pi, const float* pts_xyz, int P, dfu_stream stream) {
    DFU_REQUIRE(pi && pts_xyz, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(P >= 1, DFU_ERR_INVALID, "P must be >= 1");
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != pi->device) DFU_CUDA_OK(cudaSetDevice(pi->device));
    (void) cudaGetLastError();
    cudaStream_t st = as_stream(stream);
    if (!pi->desc) {
        DFU_CUDA_OK(cudaMalloc(&pi->desc, sizeof(PGrid)));
        DFU_CUDA_OK(cudaMalloc(&pi->cell_start, ((size_t) PG_MAX_CELLS + 1) * sizeof(int)));
        DFU_CUDA_OK(cudaMalloc(&pi->cell_count, (size_t) PG_MAX_CELLS * sizeof(int)));
        DFU_CUDA_OK(cudaMalloc(&pi->cursor, (8 + (size_t) PG_MAX_CHUNKS) * sizeof(int)));  // bbox words + chunk sums
    }
    if (P > pi->capacity) {
        cudaFree(pi->sorted);
        cudaFree(pi->occ);
        pi->sorted = nullptr;
        pi->occ = nullptr;
        pi->capacity = 0;
        DFU_CUDA_OK(cudaMalloc(&pi->sorted, (size_t) P * sizeof(float4)));
        DFU_CUDA_OK(cudaMalloc(&pi->occ, (size_t) P * sizeof(int)));
        pi->capacity = P;
    }
    pi->P = P;
    unsigned* bbox = reinterpret_cast<unsigned*>(pi->cursor);
    int* chunk_sum = pi->cursor + 8;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pi->device);
    // min words start at 0xffffffff, max words and the done counter at 0
    DFU_CUDA_OK(cudaMemsetAsync(bbox, 0xff, 3 * sizeof(unsigned), st));
    DFU_CUDA_OK(cudaMemsetAsync(bbox + 3, 0, 5 * sizeof(unsigned), st));
    pg_bbox_kernel<<<std::min(sms, div_up(P, 256)), 256, 0, st>>>(pts_xyz, P, bbox, pi->desc);
    DFU_LAUNCH_OK();
    // pass 0 measures the occupancy of the first-guess grid, pass 1 bins with the corrected cell size
    pg_zero_kernel<<<sms * 4, 256, 0, st>>>(pi->desc, pi->cell_count);
    DFU_LAUNCH_OK();
    pg_count_kernel<<<div_up(P, 256), 256, 0, st>>>(pts_xyz, P, pi->desc, pi->cell_count);
    DFU_LAUNCH_OK();
    pg_occupancy_kernel<<<sms * 4, 256, 0, st>>>(pi->desc, pi->cell_count);
    DFU_LAUNCH_OK();
    pg_adjust_kernel<<<1, 32, 0, st>>>(pi->desc);
    DFU_LAUNCH_OK();
    pg_zero_kernel<<<sms * 4, 256, 0, st>>>(pi->desc, pi->cell_count);
    DFU_LAUNCH_OK();
    pg_count_kernel<<<div_up(P, 256), 256, 0, st>>>(pts_xyz, P, pi->desc, pi->cell_count);
    DFU_LAUNCH_OK();
    pg_chunk_sum_kernel<<<PG_MAX_CHUNKS, 256, 0, st>>>(pi->desc, pi->cell_count, chunk_sum);
    DFU_LAUNCH_OK();
    pg_chunk_scan_kernel<<<1, 1024, 0, st>>>(pi->desc, chunk_sum);
    DFU_LAUNCH_OK();
    pg_offsets_kernel<<<PG_MAX_CHUNKS, 256, 0, st>>>(pi->desc, pi->cell_count, chunk_sum, pi->cell_start, pi->occ);
    DFU_LAUNCH_OK();
    pg_fill_kernel<<<div_up(P, 256), 256, 0, st>>>(pts_xyz, P, pi->desc, pi->cell_start, pi->cell_count, pi->sorted);
    DFU_LAUNCH_OK();
    pi->built = true;
    if (prev != pi->device) cudaSetDevice(prev);
    return DFU_OK;
}

int dfu_pointindex_nearest(const dfu_pointindex* pi, const float* q_xyz, int Q, int32_t* idx, float* dist2, dfu_stream stream) {
    DFU_REQUIRE(pi && (Q == 0 || (q_xyz && idx)), DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(pi->built, DFU_ERR_NOT_INIT, "point index not built");
    if (Q == 0) return DFU_OK;
    (void) cudaGetLastError();
    pg_nearest_kernel<<<div_up(Q, 128), 128, 0, as_stream(stream)>>>(q_xyz, Q, pi->desc, pi->cell_start, pi->sorted, pi->occ, idx, dist2);
    DFU_LAUNCH_OK();
    return DFU_OK;
}

int dfu_find_corresponding(dfu_pointindex* pi, const float* canon_v, const float* canon_n, int P_canon, const float* live_v, int P_live,
                           float* out_v, float* out_n, int32_t* idx_out, dfu_stream stream) {
    DFU_REQUIRE(pi && canon_v && live_v && out_v, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(P_canon >= 1 && P_live >= 0, DFU_ERR_INVALID, "bad sizes");
    int rc = dfu_pointindex_build(pi, canon_v, P_canon, stream);  // the reference rebuilds its KD-tree every call too (:221-226)
    if (rc != DFU_OK) return rc;
    if (P_live == 0) return DFU_OK;
    int32_t* idx = idx_out;
    if (!idx) {
        if (P_live > pi->tmp_cap) {
            cudaFree(pi->tmp_idx);
            pi->tmp_idx = nullptr;
            pi->tmp_cap = 0;
            DFU_CUDA_OK(cudaMalloc(&pi->tmp_idx, (size_t) P_live * sizeof(int32_t)));
            pi->tmp_cap = P_live;
        }
        idx = pi->tmp_idx;
    }
    rc = dfu_pointindex_nearest(pi, live_v, P_live, idx, nullptr, stream);
    if (rc != DFU_OK) return rc;
    gather_rows_kernel<<<div_up(P_live, 256), 256, 0, as_stream(stream)>>>(canon_v, canon_n, idx, P_live, out_v, out_n);
    DFU_LAUNCH_OK();
    return DFU_OK;
}

}  // extern "C"
