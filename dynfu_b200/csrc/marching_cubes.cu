// Marching cubes over the TSDF volume: the vertex producer of DynFusion::operator() / init
// (src/dynfu/dyn_fusion.cpp:74-88,120-134 -> kfusion::cuda::MarchingCubes::run, src/kfusion/marching_cubes.cpp:20-63 ->
//  getOccupiedVoxels / computeOffsetsAndTotalVertices / generateTriangles, src/kfusion/cuda/marching_cubes.cu:77-296).
//
// Same per-cube arithmetic as the reference (cube index from the 8 corner signs, no triangles where a corner has weight 0;
// cell centres (i + 0.5) * cell; edge vertex p0 + t (p1 - p0), t = (iso - f0) / (f1 - f0 + 1e-15f); triangle table order) in
// the canonical IEEE arithmetic of DESIGN.md.  What is different by design:
//   * any volume size (the reference hard-codes 128^3: include/kfusion/internal.hpp:74, marching_cubes.cu:151-152,283-285);
//   * a FIXED output order -- tiles of 32 x 8 x 8 cubes ascending, cubes ascending inside a tile -- instead of the order in
//     which warps win an atomic (marching_cubes.cu:108): bit-reproducible, and a valid outcome of the reference's race;
//   * three passes without a host round trip: per-tile vertex counts (the only pass that sweeps the volume, 128-bit
//     loads), a two-level scan of the tile counts, and emission that revisits only the tiles with triangles (a few percent);
//     the reference downloads two scalars between its kernels and scans with Thrust.
#include "dfu_internal.h"
#include "dfu_math.cuh"
#include "mc_tables.inc"

using namespace dfu;

namespace {

constexpr int MC_CHUNK = 1024;  // tile counts scanned by one CTA

struct McArgs {
    const uint32_t* vol;
    int dx, dy, dz;
    float cx, cy, cz;  // cell size (metres)
    int ntx, nty, ntz, ntiles;
    int* tile_count;   // per tile: vertices it emits; after the scan: its exclusive offset
    int* chunk_sum;    // per MC_CHUNK tiles
    int nchunks;
    float4* verts;
    int32_t* cube_ids;
    long capacity;
    int* n_vertices;
};

__constant__ signed char c_tri[256][16];
__constant__ unsigned char c_nverts[256];

// corner data of the four x-consecutive cubes (x .. x+3, y, z) of one thread: field and weight of the 5 x 2 x 2 voxels
struct Quad {
    float f[2][2][5];
    bool ok[2][2][5];  // weight != 0
};

DFU_DEV void load_quad(const McArgs& a, int x, int y, int z, Quad& q) {
    const size_t plane = (size_t) a.dx * a.dy;
#pragma unroll
    for (int kz = 0; kz < 2; ++kz)
#pragma unroll
        for (int ky = 0; ky < 2; ++ky) {
            const bool row_in = y + ky < a.dy && z + kz < a.dz;
            const uint32_t* row = a.vol + (size_t) x + (size_t) (y + ky) * a.dx + plane * (size_t) (z + kz);
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            uint32_t e = 0u;
            if (row_in) {
                v = __ldg(reinterpret_cast<const uint4*>(row));
                if (x + 4 < a.dx) e = __ldg(row + 4);
            }
            const uint32_t w[5] = {v.x, v.y, v.z, v.w, e};
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                q.f[kz][ky][i] = __half2float(__ushort_as_half((unsigned short) (w[i] & 0xffffu)));
                q.ok[kz][ky][i] = (w[i] >> 16) != 0u;
            }
        }
}

// cube index of cube i (0..3) of the quad, 0 when a corner has no weight (CubeIndexEstimator::computeCubeIndex, :41-75)
DFU_DEV int cube_index(const Quad& q, int i, float (&f)[8]) {
    // corners 0 (x,y,z) 1 (x+1,y,z) 2 (x+1,y+1,z) 3 (x,y+1,z) 4..7 the same at z+1
    f[0] = q.f[0][0][i]; f[1] = q.f[0][0][i + 1]; f[2] = q.f[0][1][i + 1]; f[3] = q.f[0][1][i];
    f[4] = q.f[1][0][i]; f[5] = q.f[1][0][i + 1]; f[6] = q.f[1][1][i + 1]; f[7] = q.f[1][1][i];
    const bool ok = q.ok[0][0][i] && q.ok[0][0][i + 1] && q.ok[0][1][i + 1] && q.ok[0][1][i] && q.ok[1][0][i] && q.ok[1][0][i + 1] &&
                    q.ok[1][1][i + 1] && q.ok[1][1][i];
    if (!ok) return 0;
    int ci = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) ci |= (f[c] < 0.f) ? (1 << c) : 0;
    return ci;
}

DFU_DEV void tile_origin(const McArgs& a, int tile, int& x0, int& y0, int& z0) {
    x0 = (tile % a.ntx) * 32;
    tile /= a.ntx;
    y0 = (tile % a.nty) * 8;
    z0 = (tile / a.nty) * 8;
}

// pass 1: vertices per tile
__global__ void __launch_bounds__(128) mc_count_kernel(const __grid_constant__ McArgs a) {
    __shared__ int s_sum[4];
    int x0, y0, z0;
    tile_origin(a, blockIdx.x, x0, y0, z0);
    int cnt = 0;
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        const int lin = it * 128 + threadIdx.x;
        const int x = x0 + (lin & 7) * 4, y = y0 + ((lin >> 3) & 7), z = z0 + (lin >> 6);
        if (x >= a.dx || y + 1 >= a.dy || z + 1 >= a.dz) continue;
        Quad q;
        load_quad(a, x, y, z, q);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (x + i + 1 >= a.dx) continue;
            float f[8];
            cnt += c_nverts[cube_index(q, i, f)];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) a.tile_count[blockIdx.x] = s_sum[0] + s_sum[1] + s_sum[2] + s_sum[3];
}

// pass 2a: exclusive scan of MC_CHUNK tile counts per CTA (256 threads x 4), chunk totals out
__global__ void __launch_bounds__(256) mc_scan_chunks_kernel(const __grid_constant__ McArgs a) {
    __shared__ int s_w[8];
    const int base = blockIdx.x * MC_CHUNK + threadIdx.x * 4;
    int v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = base + i < a.ntiles ? a.tile_count[base + i] : 0;
    const int mine = v[0] + v[1] + v[2] + v[3];
    int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < (int) (threadIdx.x >> 5); ++w) wbase += s_w[w];
    int run = wbase + inc - mine;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (base + i < a.ntiles) a.tile_count[base + i] = run;
        run += v[i];
    }
    if (threadIdx.x == 255) a.chunk_sum[blockIdx.x] = run;
}

// pass 2b: exclusive scan of the chunk totals (one CTA), total vertex count out
__global__ void __launch_bounds__(1024) mc_scan_totals_kernel(const __grid_constant__ McArgs a) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < a.nchunks; c0 += 1024) {
        const int i = c0 + threadIdx.x;
        const int v = i < a.nchunks ? a.chunk_sum[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((threadIdx.x & 31) >= o) inc += t;
        }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
        __syncthreads();
        int wbase = s_carry;
        for (int w = 0; w < (int) (threadIdx.x >> 5); ++w) wbase += s_w[w];
        if (i < a.nchunks) a.chunk_sum[i] = wbase + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = wbase + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *a.n_vertices = s_carry;
}

// pass 3: emission, tiles with triangles only
__global__ void __launch_bounds__(128) mc_emit_kernel(const __grid_constant__ McArgs a) {
    __shared__ int s_w[4];
    __shared__ int s_run;
    const int tile = blockIdx.x;
    const int begin = a.tile_count[tile] + a.chunk_sum[tile / MC_CHUNK];
    const int next = tile + 1 < a.ntiles ? a.tile_count[tile + 1] + a.chunk_sum[(tile + 1) / MC_CHUNK] : *a.n_vertices;
    if (next == begin) return;  // (uniform over the CTA)
    int x0, y0, z0;
    tile_origin(a, tile, x0, y0, z0);
    if (threadIdx.x == 0) s_run = begin;
    __syncthreads();
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        const int lin = it * 128 + threadIdx.x;
        const int x = x0 + (lin & 7) * 4, y = y0 + ((lin >> 3) & 7), z = z0 + (lin >> 6);
        Quad q;
        int ci[4] = {0, 0, 0, 0};
        int cnt = 0;
        const bool in = x < a.dx && y + 1 < a.dy && z + 1 < a.dz;
        if (in) {
            load_quad(a, x, y, z, q);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float f[8];
                ci[i] = x + i + 1 < a.dx ? cube_index(q, i, f) : 0;
                cnt += c_nverts[ci[i]];
            }
        }
        // exclusive scan of the per-thread counts in thread order (= ascending cube order inside the tile)
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((threadIdx.x & 31) >= o) inc += t;
        }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
        __syncthreads();
        int off = s_run + inc - cnt;
        for (int w = 0; w < (int) (threadIdx.x >> 5); ++w) off += s_w[w];
        const int total = s_w[0] + s_w[1] + s_w[2] + s_w[3];
        __syncthreads();
        if (threadIdx.x == 0) s_run += total;
        if (cnt > 0) {
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
                const int nv = c_nverts[ci[i]];
                if (nv == 0) continue;
                float f[8];
                cube_index(q, i, f);
                const int cube = (x + i) + a.dx * (y + a.dy * z);
                for (int k = 0; k < nv; ++k) {
                    const int e = c_tri[ci[i]][k];
                    // edges 0-3: bottom ring (c, c+1 mod 4), 4-7: top ring, 8-11: verticals (c, c+4)
                    const int ca = e < 8 ? e : e - 8;
                    const int cb = e < 4 ? (e + 1) & 3 : (e < 8 ? 4 + ((e + 1) & 3) : e - 4);
                    const int ax = ((ca & 3) == 1 || (ca & 3) == 2) ? 1 : 0, ay = (ca & 3) >= 2 ? 1 : 0, az = ca >> 2;
                    const int bx = ((cb & 3) == 1 || (cb & 3) == 2) ? 1 : 0, by = (cb & 3) >= 2 ? 1 : 0, bz = cb >> 2;
                    // getNodeCoo (:181-190): (index + 0.5) * cell
                    const float pax = fmul(fadd((float) (x + i + ax), 0.5f), a.cx), pay = fmul(fadd((float) (y + ay), 0.5f), a.cy),
                                paz = fmul(fadd((float) (z + az), 0.5f), a.cz);
                    const float pbx = fmul(fadd((float) (x + i + bx), 0.5f), a.cx), pby = fmul(fadd((float) (y + by), 0.5f), a.cy),
                                pbz = fmul(fadd((float) (z + bz), 0.5f), a.cz);
                    // vertex_interp (:192-199)
                    const float t = __fdiv_rn(fsub(0.f, f[ca]), fadd(fsub(f[cb], f[ca]), 1e-15f));
                    const long o = (long) off + k;
                    if (o < a.capacity) {
                        a.verts[o] = make_float4(fadd(pax, fmul(t, fsub(pbx, pax))), fadd(pay, fmul(t, fsub(pby, pay))),
                                                 fadd(paz, fmul(t, fsub(pbz, paz))), 1.f);  // store_point (:262-264)
                        if (a.cube_ids) a.cube_ids[o] = cube;
                    }
                }
                off += nv;
            }
        }
        __syncthreads();
    }
}

int upload_tables(int device) {
    static bool done[64] = {};
    if (device >= 0 && device < 64 && done[device]) return DFU_OK;
    unsigned char nv[256];
    for (int i = 0; i < 256; ++i) {
        int n = 0;
        while (n < 16 && MC_TRI_TABLE[i][n] >= 0) ++n;
        nv[i] = (unsigned char) n;
    }
    DFU_CUDA_OK(cudaMemcpyToSymbol(c_tri, MC_TRI_TABLE, sizeof(MC_TRI_TABLE)));
    DFU_CUDA_OK(cudaMemcpyToSymbol(c_nverts, nv, sizeof(nv)));
    if (device >= 0 && device < 64) done[device] = true;
    return DFU_OK;
}

}  // namespace

extern "C" int dfu_marching_cubes(const void* volume, const int dims[3], const float volume_size[3], void* vertices_xyz1,
                                  int32_t* cube_ids, long capacity, int* n_vertices_dev, dfu_stream stream) {
    DFU_REQUIRE(volume && dims && volume_size && n_vertices_dev, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(vertices_xyz1 || capacity == 0, DFU_ERR_INVALID, "no output buffer");
    DFU_REQUIRE(dims[0] > 0 && dims[0] % 4 == 0 && dims[1] > 0 && dims[2] > 0, DFU_ERR_INVALID, "dims.x must be a positive multiple of 4");
    DFU_REQUIRE(((uintptr_t) volume & 15) == 0, DFU_ERR_INVALID, "volume must be 16-byte aligned");
    DFU_REQUIRE(capacity >= 0, DFU_ERR_INVALID, "negative capacity");
    DFU_GUARD(dfu_device_of(volume));
    cudaStream_t st = as_stream(stream);
    int device = 0;
    DFU_CUDA_OK(cudaGetDevice(&device));
    int rc = upload_tables(device);
    if (rc != DFU_OK) return rc;
    McArgs a{};
    a.vol = reinterpret_cast<const uint32_t*>(volume);
    a.dx = dims[0]; a.dy = dims[1]; a.dz = dims[2];
    a.cx = volume_size[0] / (float) dims[0];  // generateTriangles :283-285 (volume_size / cells)
    a.cy = volume_size[1] / (float) dims[1];
    a.cz = volume_size[2] / (float) dims[2];
    a.ntx = div_up(dims[0] - 1, 32);
    a.nty = div_up(dims[1] - 1, 8);
    a.ntz = div_up(dims[2] - 1, 8);
    const long nt = (long) a.ntx * a.nty * a.ntz;
    if (nt <= 0) {
        DFU_CUDA_OK(cudaMemsetAsync(n_vertices_dev, 0, sizeof(int), st));
        return DFU_OK;
    }
    DFU_REQUIRE(nt <= 0x7fffffffL, DFU_ERR_INVALID, "volume too large for one launch");
    a.ntiles = (int) nt;
    a.nchunks = div_up(nt, MC_CHUNK);
    a.verts = reinterpret_cast<float4*>(vertices_xyz1);
    a.cube_ids = cube_ids;
    a.capacity = capacity;
    a.n_vertices = n_vertices_dev;
    char* scratch = nullptr;
    const size_t bytes = ((size_t) nt + (size_t) a.nchunks + 64) * sizeof(int);
    DFU_CUDA_OK(scratch_alloc((void**) &scratch, bytes, st));
    a.tile_count = reinterpret_cast<int*>(scratch);
    a.chunk_sum = a.tile_count + nt;
    mc_count_kernel<<<a.ntiles, 128, 0, st>>>(a);
    ++g_dfu_launches;
    mc_scan_chunks_kernel<<<a.nchunks, 256, 0, st>>>(a);
    ++g_dfu_launches;
    mc_scan_totals_kernel<<<1, 1024, 0, st>>>(a);
    ++g_dfu_launches;
    mc_emit_kernel<<<a.ntiles, 128, 0, st>>>(a);
    ++g_dfu_launches;
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    if (e != cudaSuccess) {
        dfu_set_error("dfu_marching_cubes: %s", cudaGetErrorString(e));
        return DFU_ERR_CUDA;
    }
    return DFU_OK;
}
