// Small device-side building blocks shared by frontend.cu and update.cu: single-CTA exclusive scan, ordered-uint
// float encoding for atomic min/max, ordered (stable) stream compaction by warp ballots.
#pragma once
#include "dfu_math.cuh"

namespace {

// exclusive scan of n counts (single CTA), total to *total
__global__ void __launch_bounds__(1024) small_scan_kernel(int* __restrict__ counts, int n, int* __restrict__ total) {
    __shared__ int sh[1024];
    const int per = (n + 1023) / 1024;
    const int lo = min(n, (int) threadIdx.x * per), hi = min(n, lo + per);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += counts[i];
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
        const int c = counts[i];
        counts[i] = run;
        run += c;
    }
    if (threadIdx.x == 1023) *total = sh[1023];
}

// order-preserving float <-> uint map for atomicMin/Max
DFU_DEV unsigned f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
DFU_DEV float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }


}  // namespace
