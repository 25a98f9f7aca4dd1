// Internal declarations shared by the translation units of libdynfu_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/dynfu_b200.h"

// ---- error plumbing: status codes out, message kept per thread (no exceptions, no exit()) ----------
void dfu_set_error(const char* fmt, ...);
#define DFU_CUDA_OK(expr)                                                                         \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            dfu_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));  \
            return DFU_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)
#define DFU_REQUIRE(cond, code, msg)                               \
    do {                                                           \
        if (!(cond)) {                                             \
            dfu_set_error("%s: %s", __func__, msg);                \
            return code;                                           \
        }                                                          \
    } while (0)
// every kernel launch of the library goes through this macro, so the counter is the number of OUR kernels launched
extern unsigned long long g_dfu_launches;
#define DFU_LAUNCH_OK()                     \
    do {                                    \
        ++g_dfu_launches;                   \
        DFU_CUDA_OK(cudaGetLastError());    \
    } while (0)

// ---- device selection: every entry point runs on the device that owns its handle / buffers, whatever the caller's current
// device is, and puts the caller's device back on every return path
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};
// also drops any stale, non-sticky error another library left in the runtime, so that the launch checks
// below report only this library's own failures
#define DFU_GUARD(dev)                                                     \
    DeviceGuard _guard(dev);                                               \
    if (!_guard.ok) {                                                      \
        dfu_set_error("%s: cannot select CUDA device %d", __func__, dev);  \
        return DFU_ERR_CUDA;                                               \
    }                                                                      \
    (void) cudaGetLastError();
// device that owns a device pointer (the current device when the runtime cannot tell)
static inline int dfu_device_of(const void* p) {
    cudaPointerAttributes at;
    int cur = 0;
    cudaGetDevice(&cur);
    if (p && cudaPointerGetAttributes(&at, p) == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged))
        return at.device;
    (void) cudaGetLastError();
    return cur;
}

static inline cudaStream_t as_stream(dfu_stream s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int div_up(long a, long b) { return (int) ((a + b - 1) / b); }
cudaMemPool_t scratch_pool(int device);  // tsdf.cu: private stream-ordered pool for per-call scratch
static inline cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t st) {
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return e;
    cudaMemPool_t pool = scratch_pool(device);
    return pool ? cudaMallocFromPoolAsync(p, bytes, pool, st) : cudaMallocAsync(p, bytes, st);
}

// ---- the warp field handle ---------------------------------------------------------------------------
// Node state in HBM (DESIGN.md "data layout"): three float4 arrays padded to a multiple of 32 entries.
//   pos_w[i] = (dg_v.x, dg_v.y, dg_v.z, dg_w)   padding: (+inf,+inf,+inf,1)  -> never a neighbour
//   real[i], dual[i]                             dg_se3 as two quaternions (w,x,y,z)
struct BrickTable {            // per-volume-geometry acceleration data for the warped integrator
    float2* bounds = nullptr;  // per 8x8x8 brick: (squared dist of brick centre to its 8th, and to its 1st node)
    size_t capacity = 0;       // bricks allocated
    int dims[3] = {0, 0, 0};
    float voxel[3] = {0, 0, 0};
    uint64_t node_epoch = 0;   // positions epoch the bounds were built for
    uint64_t cache_epoch = 0;  // positions epoch the built[] flags of the voxel cache are valid for
    bool valid = false;
    // per-voxel 8-NN + weight cache (filled lazily by the integrator, valid while node POSITIONS are unchanged: ids
    // and Gaussian weights depend only on the canonical voxel and node positions, never on the node transforms):
    // per brick 512 voxels x (8 u16 ids = 16 B | 8 f32 weights = 32 B) = 24 KB; built[brick] != 0 once filled.
    // The pools cover the brick planes [pool_zb0, pool_zb1) (the z-slab this rank integrates).
    uint4* knn_pool = nullptr;
    float4* w_pool = nullptr;
    unsigned char* built = nullptr;
    size_t pool_bricks = 0;    // bricks the pools have room for (0: cache disabled)
    int pool_zb0 = 0, pool_zb1 = 0;
};

// Uniform grid over the node positions for the point-query kNN (rebuilt when positions change)
struct GridDesc {
    float ox, oy, oz;  // origin (bbox min)
    float h, inv_h;    // cell edge
    int nx, ny, nz;
    int n_occ;         // number of non-empty cells (entries of NodeGrid::occ)
};
constexpr int DFU_GRID_MAX_DIM = 128;
constexpr int DFU_GRID_MAX_CELLS = DFU_GRID_MAX_DIM * DFU_GRID_MAX_DIM * DFU_GRID_MAX_DIM;
struct NodeGrid {
    GridDesc* desc = nullptr;     // device
    int* cell_start = nullptr;    // DFU_GRID_MAX_CELLS + 1
    int* cell_count = nullptr;    // DFU_GRID_MAX_CELLS (scratch during the build)
    float4* sorted = nullptr;     // nodes ordered by cell: (x, y, z, index bits)
    int* occ = nullptr;           // ids of the non-empty cells (at most N)
    int capacity = 0;             // entries of `sorted`
    uint64_t node_epoch = 0;
    bool valid = false;
};

struct dfu_warpfield {
    int device = 0;
    int N = 0;            // nodes
    int Npad = 0;         // padded to 32
    int capacity = 0;     // allocated entries
    float epsilon = 0.f;
    float4* pos_w = nullptr;
    float4* real = nullptr;
    float4* dual = nullptr;
    // device flags/scalars: [0] = 1 if every node has real == (1,0,0,0) and dual.w == 0 (translation-only field),
    //                       [1] = max dg_w as float bits
    int* flags = nullptr;
    float* staging = nullptr;  // device staging for *_host uploads (N*12 floats)
    size_t staging_cap = 0;
    uint64_t node_epoch = 0;   // bumped when positions change
    BrickTable bricks;
    NodeGrid grid;
    bool initialised = false;
};

// kernels living in warpfield.cu that tsdf.cu / solver.cu launch
int dfu_wf_refresh_flags(dfu_warpfield* wf, cudaStream_t st);
int dfu_wf_build_brick_table(dfu_warpfield* wf, const int dims[3], const float voxel[3], int z0, int z1, cudaStream_t st);
int dfu_wf_build_data_graph(const dfu_warpfield* wf, const float* canon, const float* live, int P, int32_t* nbr,
                            float* wts, float* dvec, int* deg, int32_t* rank, cudaStream_t st);
int dfu_wf_build_node_graph(const dfu_warpfield* wf, int32_t* nnbr, cudaStream_t st);
