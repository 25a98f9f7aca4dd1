// oracle/ref_shim: CUDA texture *references* (removed in CUDA 12) restated as plain __device__ descriptors, so that the
// reference's `texture<T, dim> name;` declarations, tex2D / tex1Dfetch reads and cudaBindTexture calls compile unchanged.
//   tsdf_volume.cu:41,108-112  2-D, cudaFilterModePoint, cudaAddressModeBorder, unnormalised coordinates, half channel
//                              read as float: texel (floor(x), floor(y)), 0 outside the image;
//   marching_cubes.cu:9-22     1-D int fetches of the triangle / vertex-count tables.
// Included through the cudaUtil.h shim (every reference .cu reaches it via temp_utils.hpp:4).
// Test infrastructure only (oracle/_ref); nothing here is linked into the product.
#pragma once
#include <cuda_runtime.h>

template <class T, int Dim, int Mode = 0>
struct ref_shim_texture {
    const void *data = nullptr;
    int cols = 0, rows = 0;
    size_t step = 0;
    int filterMode = 0;
    int addressMode[3] = {0, 0, 0};
    template <class... A>
    __host__ __device__ constexpr ref_shim_texture(A...) {}
    __host__ __device__ constexpr ref_shim_texture() {}
};
struct ref_shim_binding {
    const void *data;
    int cols, rows;
    size_t step;
};
#define texture __device__ ref_shim_texture
#define cudaCreateChannelDescHalf() 0

// point-sampled, border-addressed half texel as float (the only 2-D texture the reference declares)
__device__ __forceinline__ float tex2D(const ref_shim_texture<float, 2, 0> &t, float x, float y) {
    const int ix = (int) floorf(x), iy = (int) floorf(y);
    if (ix < 0 || iy < 0 || ix >= t.cols || iy >= t.rows) return 0.f;
    const unsigned short h = *reinterpret_cast<const unsigned short *>(static_cast<const char *>(t.data) + iy * t.step + ix * 2);
    return __half2float(h);
}
__device__ __forceinline__ int tex1Dfetch(const ref_shim_texture<int, 1, 0> &t, int i) {
    return static_cast<const int *>(t.data)[i];
}


// cudaBindTexture / cudaUnbindTexture for the 1-D int tables of marching_cubes.cu:32-52
template <class T>
inline cudaError_t cudaBindTexture(size_t *, const ref_shim_texture<T, 1, 0> &tex, const void *ptr, const cudaChannelFormatDesc &,
                                   size_t bytes = 0) {
    ref_shim_binding b{ptr, (int) (bytes / sizeof(T)), 1, 0};
    return cudaMemcpyToSymbol(tex, &b, sizeof(b));
}
template <class T>
inline cudaError_t cudaUnbindTexture(const ref_shim_texture<T, 1, 0> &) {
    return cudaSuccess;
}
