// Microbenchmarks behind the C-ABI: the FP32 denominator of the roofline (MEASURED_PEAKS.json has
// the HBM copy rate and a bf16 GEMM, not the FP32 SIMT peak the kNN kernels are bound by).
#include "dfu_internal.h"

namespace {
// 16 independent FFMA chains per thread (3-register form, nothing for the compiler to fold), 8 CTAs of 256 threads per SM
constexpr int MB_CHAINS = 16;
__global__ void __launch_bounds__(256) fma_chain_kernel(float* out, int iters, float a, float b) {
    float x[MB_CHAINS];
#pragma unroll
    for (int i = 0; i < MB_CHAINS; ++i) x[i] = (float) (threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < MB_CHAINS; ++i) x[i] = __fmaf_rn(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MB_CHAINS; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;  // never true: keeps the chains alive
}
}  // namespace

extern "C" int dfu_microbench_fp32(int device, double* tflops_out, double* sm_mhz_out) {
    DFU_REQUIRE(tflops_out != nullptr, DFU_ERR_INVALID, "null output");
    int prev = 0;
    DFU_CUDA_OK(cudaGetDevice(&prev));
    DFU_CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    DFU_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    float* out = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = DFU_OK;
    double best = 0.0;
    const int iters = 8192, grid = prop.multiProcessorCount * 8;
    if (cudaMalloc(&out, sizeof(float)) != cudaSuccess || cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
        rc = DFU_ERR_CUDA;
    } else {
        for (int rep = 0; rep < 6 && rc == DFU_OK; ++rep) {  // first repetitions warm the clocks up; best of the rest
            cudaEventRecord(e0);
            fma_chain_kernel<<<grid, 256>>>(out, iters, 0.999f, 1e-3f);
            ++g_dfu_launches;
            cudaEventRecord(e1);
            if (cudaEventSynchronize(e1) != cudaSuccess) {
                rc = DFU_ERR_CUDA;
                break;
            }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            const double fl = 2.0 * MB_CHAINS * (double) iters * 256.0 * grid;
            if (rep >= 2 && ms > 0.f) best = fl / (ms * 1e-3) / 1e12 > best ? fl / (ms * 1e-3) / 1e12 : best;
        }
    }
    if (rc != DFU_OK) dfu_set_error("dfu_microbench_fp32: %s", cudaGetErrorString(cudaGetLastError()));
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(out);
    cudaSetDevice(prev);
    *tflops_out = best;
    // clock implied by the measurement: flops / (SMs * 128 lanes * 2)
    if (sm_mhz_out) *sm_mhz_out = best * 1e12 / (prop.multiProcessorCount * 128.0 * 2.0) / 1e6;
    return rc;
}
