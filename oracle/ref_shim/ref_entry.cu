// oracle/ref_shim/ref_entry.cu -- TEST INFRASTRUCTURE ONLY.
//
// A C-ABI over the reference's OWN CUDA kernels, which oracle/Makefile compiles where they lie under /root/reference
// (src/kfusion/cuda/{tsdf_volume,imgproc,marching_cubes}.cu + src/kfusion/device_memory.cpp, with the reference's nvcc
// flags CMakeLists.txt:76-78) into oracle/_ref/libdynfu_ref_cuda.so.  Only `-m gpu` tests load it; it pins the
// restatements in oracle/dynfu_oracle.cpp and the product's kernels against the thing they replace:
//   ref_integrate        device::integrate            src/kfusion/cuda/tsdf_volume.cu:100-121 (TsdfIntegrator :43-94)
//   ref_clear_volume     device::clear_volume         tsdf_volume.cu:26-34
//   ref_compute_dists    device::compute_dists        src/kfusion/cuda/imgproc.cu:248-254
//   ref_points_normals   device::computePointNormals  imgproc.cu:218-226
//   ref_raycast_points   device::raycast (points)     tsdf_volume.cu:372-386 (TsdfRaycaster :173-337)
//   ref_marching_cubes   getOccupiedVoxels / computeOffsetsAndTotalVertices / generateTriangles
//                                                     src/kfusion/cuda/marching_cubes.cu:144-161,163-179,268-296
//   ref_mc_tables        the tables of src/kfusion/marching_cubes.cpp:66-354
// All pointers are device pointers unless named *_host.  Nothing of this file or of oracle/_ref is linked into, loaded by
// or shipped with the product library.
#include <kfusion/cuda/device.hpp>

#include <cstdio>

extern const int edgeTable[256];
extern const int triTable[256][16];
extern const int numVertsTable[256];

// the three one-line constructors the reference defines in src/kfusion/precomp.cpp:22,43,57 (that file also holds
// kfusion::Intr, which needs OpenCV, so it cannot be compiled here)
kfusion::device::TsdfVolume::TsdfVolume(elem_type *data, int3 dims, float3 voxel_size, float trunc_dist, int max_weight)
    : data(data), dims(dims), voxel_size(voxel_size), trunc_dist(trunc_dist), max_weight(max_weight) {}
kfusion::device::Projector::Projector(float fx, float fy, float cx, float cy) : f(make_float2(fx, fy)), c(make_float2(cx, cy)) {}
kfusion::device::Reprojector::Reprojector(float fx, float fy, float cx, float cy)
    : finv(make_float2(1.f / fx, 1.f / fy)), c(make_float2(cx, cy)) {}

namespace {
using namespace kfusion;
device::TsdfVolume make_volume(void *vol, const int *dims, const float *voxel, float trunc, int max_weight) {
    return device::TsdfVolume(static_cast<ushort2 *>(vol), make_int3(dims[0], dims[1], dims[2]),
                              make_float3(voxel[0], voxel[1], voxel[2]), trunc, max_weight);
}
device::Aff3f make_aff(const float *a) {  // 9 floats row-major R, then t
    device::Aff3f r;
    for (int i = 0; i < 3; ++i) r.R.data[i] = make_float3(a[3 * i], a[3 * i + 1], a[3 * i + 2]);
    r.t = make_float3(a[9], a[10], a[11]);
    return r;
}
int status() {
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        std::fprintf(stderr, "ref_entry: %s\n", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
}  // namespace

extern "C" {

int ref_clear_volume(void *vol, const int *dims_host) {
    const float one[3] = {1.f, 1.f, 1.f};
    device::clear_volume(make_volume(vol, dims_host, one, 1.f, 1));
    return status();
}

int ref_integrate(void *vol, const int *dims_host, const float *voxel_host, float trunc, int max_weight, const void *dists,
                  size_t dists_step, int rows, int cols, const float *vol2cam_host, const float *intr_host) {
    device::TsdfVolume v = make_volume(vol, dims_host, voxel_host, trunc, max_weight);
    device::Dists d;
    d.data = static_cast<unsigned short *>(const_cast<void *>(dists));
    d.step = dists_step;
    d.rows = rows;
    d.cols = cols;
    device::integrate(d, v, make_aff(vol2cam_host), device::Projector(intr_host[0], intr_host[1], intr_host[2], intr_host[3]));
    return status();
}

int ref_compute_dists(const void *depth, size_t depth_step, void *dists, size_t dists_step, int rows, int cols,
                      const float *intr_host) {
    device::Depth dp(rows, cols, const_cast<void *>(depth), depth_step);
    device::Dists d;
    d.data = static_cast<unsigned short *>(dists);
    d.step = dists_step;
    d.rows = rows;
    d.cols = cols;
    device::compute_dists(dp, d, make_float2(intr_host[0], intr_host[1]), make_float2(intr_host[2], intr_host[3]));
    return status();
}

// points / normals: dense float4 [rows][cols]
int ref_points_normals(const void *depth, size_t depth_step, int rows, int cols, const float *intr_host, void *points,
                       void *normals) {
    device::Depth dp(rows, cols, const_cast<void *>(depth), depth_step);
    device::Points p(rows, cols, points, cols * sizeof(float4));
    device::Normals n(rows, cols, normals, cols * sizeof(float4));
    device::computePointNormals(device::Reprojector(intr_host[0], intr_host[1], intr_host[2], intr_host[3]), dp, p, n);
    return status();
}

int ref_raycast_points(void *vol, const int *dims_host, const float *voxel_host, float trunc, int max_weight,
                       const float *cam2vol_host, const float *rinv_host, const float *intr_host, int rows, int cols,
                       void *points, void *normals, float step_factor, float delta_factor) {
    device::TsdfVolume v = make_volume(vol, dims_host, voxel_host, trunc, max_weight);
    device::Mat3f rinv;
    for (int i = 0; i < 3; ++i) rinv.data[i] = make_float3(rinv_host[3 * i], rinv_host[3 * i + 1], rinv_host[3 * i + 2]);
    device::Points p(rows, cols, points, cols * sizeof(float4));
    device::Normals n(rows, cols, normals, cols * sizeof(float4));
    device::raycast(v, make_aff(cam2vol_host), rinv, device::Reprojector(intr_host[0], intr_host[1], intr_host[2], intr_host[3]), p,
                    n, step_factor, delta_factor);
    return status();
}

void ref_mc_tables(int *edge_host, int *tri_host, int *numverts_host) {
    for (int i = 0; i < 256; ++i) {
        edge_host[i] = edgeTable[i];
        numverts_host[i] = numVertsTable[i];
        for (int j = 0; j < 16; ++j) tri_host[16 * i + j] = triTable[i][j];
    }
}

// The reference's marching cubes (hard-coded to a 128^3 volume, internal.hpp:74, marching_cubes.cu:151-152,283-285).
// occupied: int[3][max_voxels] scratch (voxel ids | vertex counts | offsets); triangles: float4[max_vertices].
// The order in which occupied voxels are emitted is decided by atomics (marching_cubes.cu:108) -- callers sort by voxel id.
int ref_marching_cubes(void *vol, const float *voxel_host, float trunc, int max_weight, const float *volume_size_host,
                       int *occupied, int max_voxels, void *triangles, int max_vertices, int *n_voxels_host,
                       int *n_vertices_host) {
    const int dims[3] = {128, 128, 128};
    device::TsdfVolume v = make_volume(vol, dims, voxel_host, trunc, max_weight);
    int *tabs = nullptr;
    if (cudaMalloc(&tabs, (256 + 256 * 16 + 256) * sizeof(int)) != cudaSuccess) return 1;
    cudaMemcpy(tabs, edgeTable, 256 * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(tabs + 256, &triTable[0][0], 256 * 16 * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(tabs + 256 + 4096, numVertsTable, 256 * sizeof(int), cudaMemcpyHostToDevice);
    device::bindTextures(tabs, tabs + 256, tabs + 256 + 4096);
    kfusion::cuda::DeviceArray2D<int> occ(3, max_voxels, occupied, max_voxels * sizeof(int));
    const int active = device::getOccupiedVoxels(v, occ);
    *n_voxels_host = active;
    *n_vertices_host = 0;
    if (active > 0) {
        kfusion::cuda::DeviceArray2D<int> occ_active(3, active, occupied, max_voxels * sizeof(int));
        const int total = device::computeOffsetsAndTotalVertices(occ_active);
        *n_vertices_host = total;
        if (total <= max_vertices) {
            kfusion::cuda::DeviceArray<device::PointType> out(static_cast<device::PointType *>(triangles), max_vertices);
            device::generateTriangles(v, occ_active, make_float3(volume_size_host[0], volume_size_host[1], volume_size_host[2]), out);
        }
    }
    device::unbindTextures();
    const int rc = status();
    cudaFree(tabs);
    return rc;
}

}  // extern "C"
