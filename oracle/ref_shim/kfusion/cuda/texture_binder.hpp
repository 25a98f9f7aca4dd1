// oracle/ref_shim: shadows /root/reference/include/kfusion/cuda/texture_binder.hpp.
// The reference's kernels read their inputs through CUDA *texture references* (texture<T, dim> at namespace scope,
// tex2D / tex1Dfetch, cudaBindTexture*), an API that CUDA 12 removed.  This shim keeps those source lines compiling
// unchanged: `texture` becomes a __device__ descriptor {pointer, size, pitch}, the binder fills it with
// cudaMemcpyToSymbol, and tex2D / tex1Dfetch reproduce what the reference configures --
//   tsdf_volume.cu:41,108-112  2-D, cudaFilterModePoint, cudaAddressModeBorder, unnormalised coordinates, half channel
//                              read as float: texel (floor(x), floor(y)), 0 outside the image;
//   marching_cubes.cu:9-10     1-D int fetches.
// Test infrastructure only (oracle/_ref); nothing here is linked into the product.
#pragma once
#include <cudaUtil.h>
#include <ref_shim_texture.h>
#include <cuda_runtime.h>
#include <kfusion/cuda/device_array.hpp>
#include <kfusion/safe_call.hpp>

namespace kfusion {
namespace cuda {
class TextureBinder {
public:
    // (array, texture, channel descriptor) -- tsdf_volume.cu:112
    template <class A, class T, int D, int M>
    TextureBinder(const A &arr, const ref_shim_texture<T, D, M> &tex, int /*desc*/) {
        ref_shim_binding b{arr.data, (int) arr.cols, (int) arr.rows, (size_t) arr.step};
        cudaSafeCall(cudaMemcpyToSymbol(tex, &b, sizeof(b)));
    }
};
}  // namespace cuda
namespace device {
using kfusion::cuda::TextureBinder;
}
}  // namespace kfusion

