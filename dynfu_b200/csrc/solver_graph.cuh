// Solver: transposed data graph / regularisation graph construction
// (textually included by solver.cu inside its anonymous namespace -- one translation unit)
#pragma once

// ---- graph construction ------------------------------------------------------------------------------------
__global__ void k_count(const int32_t* __restrict__ key, long n, int* __restrict__ deg) {
    const long e = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) atomicAdd(&deg[key[e]], 1);
}
// exclusive scan of deg[0..N) -> ptr[0..N], single block
__global__ void __launch_bounds__(1024) k_scan(const int* __restrict__ deg, int N, int* __restrict__ ptr) {
    __shared__ int sh[1024];
    const int per = (N + 1023) / 1024;
    const int lo = min(N, (int) threadIdx.x * per), hi = min(N, lo + per);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += deg[i];
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
        ptr[i] = run;
        run += deg[i];
    }
    if (threadIdx.x == 1023) ptr[N] = sh[1023];
}
// scatter entry ids into their node's segment (arrival order; sorted afterwards)
__global__ void k_fill(const int32_t* __restrict__ key, long n, const int* __restrict__ ptr, int* __restrict__ cursor,
                       int32_t* __restrict__ out, int shift) {
    const long e = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) {
        const int m = key[e];
        out[ptr[m] + atomicAdd(&cursor[m], 1)] = (int32_t) (e >> shift);
    }
}
// the same, writing the final (point, weight) lists directly in arrival order: versions 3 / 3r of the solver only consume
// them through order-independent (fixed-point) sums, so they skip the sort
__global__ void k_fill_emit(const int32_t* __restrict__ key, long n, const int* __restrict__ ptr, int* __restrict__ cursor,
                            const float* __restrict__ wts, int32_t* __restrict__ tv, float* __restrict__ tw) {
    const long e = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) {
        const int m = key[e];
        const int slot = ptr[m] + atomicAdd(&cursor[m], 1);
        tv[slot] = (int32_t) (e >> 3);
        tw[slot] = wts[e];
    }
}
// the same without atomics: the kNN kernel kept, per edge, the value its counting atomic returned (the edge's position in its
// node's list, in arrival order)
__global__ void k_fill_emit_ranked(const int32_t* __restrict__ key, const int32_t* __restrict__ rank, long n,
                                   const int* __restrict__ ptr, const float* __restrict__ wts, int32_t* __restrict__ tv,
                                   float* __restrict__ tw) {
    const long e = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) {
        const int m = key[e];
        if (m < 0) return;
        const int slot = ptr[m] + rank[e];
        tv[slot] = (int32_t) (e >> 3);
        tw[slot] = wts[e];
    }
}
// in-edge lists of the regularisation graph are short: insertion sort, one thread per node
__global__ void k_sort_small(const int* __restrict__ ptr, int N, int32_t* __restrict__ a) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int lo = ptr[n], hi = ptr[n + 1];
    for (int i = lo + 1; i < hi; ++i) {
        const int key = a[i];
        int j = i - 1;
        while (j >= lo && a[j] > key) {
            a[j + 1] = a[j];
            --j;
        }
        a[j + 1] = key;
    }
}
// transposed data graph: rank-sort each node's entry ids (one warp per node) and emit (point, weight) pairs
__global__ void __launch_bounds__(TPB) k_sort_emit(const int* __restrict__ ptr, int N, const int32_t* __restrict__ ent,
                                                   const float* __restrict__ wts, int32_t* __restrict__ tv,
                                                   float* __restrict__ tw) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int n = gw; n < N; n += nw) {
        const int lo = ptr[n], hi = ptr[n + 1];
        for (int i = lo + lane; i < hi; i += 32) {
            const int key = ent[i];
            int rank = 0;
            for (int j = lo; j < hi; ++j) rank += ent[j] < key;
            tv[lo + rank] = key >> 3;
            tw[lo + rank] = wts[key];
        }
    }
}

