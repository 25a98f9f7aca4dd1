// GPU Gauss-Newton / PCG solve of the warp-field energy.
//
// Replaces CombinedSolver + Opt + the Terra energy (src/dynfu/utils/opt_solver.cpp,
// include/dynfu/utils/terra/energy.t; Opt itself is not in the reference tree):
//   E(t) = sum_v tukey_v * | live_v - canon_v - sum_k w_vk t[n_vk] |^2            (energy.t:47-55)
//        + sum_n sum_i w_reg^2 * | t[m_ni] - t[n] |^2                             (energy.t:73-78)
// with w_vk = exp(-|canon_v - dg_v[n]|^2 / (2 dg_w[n]^2)) (energy.t:15-17), w_reg^2 = lambda/(N*8)
// (opt_solver.cpp:30).  The residuals are linear in t, so one GN step is one SPD solve
// (W^T Theta W + w_reg^2 L) delta = -J^T r, done matrix-free by block-Jacobi-preconditioned CG.
//
// Device layout: per point 8 neighbour ids + 8 weights (two 128-bit loads each), d = live - canon, tukey;
// per node t, delta, r, z, p, q (N*3 floats each) and the normal-equation buffer
//   nbuf = [ b = -J^T r : 3N | D = diag(J^T J) : N | E : 4 ]
// which is the ONE buffer a data-parallel run all-reduces per GN step; the PCG all-reduces q (3N floats)
// per iteration.  Per-point contributions reach the per-node blocks with float atomics (red.global.add);
// everything that follows the all-reduce (regularisation term, dot products, vector updates) is
// evaluated in a fixed order, so every rank computes bit-identical iterates.
// No host synchronisation happens inside solve_all unless early_out / pcg_tol ask for it.
#include <limits.h>

#include "dfu_internal.h"
#include "dfu_math.cuh"

using namespace dfu;

namespace {

constexpr int TPB = 256;
constexpr int MAX_PARTIALS = 1024;

struct Scalars {
    double rz[2];        // r.z of PCG iteration it is rz[it & 1]
    double rz_ref;       // r.z of the first GN step of this solve (< 0: unset)
    double E;            // energy at the last assemble (data + reg)
    double E0;           // energy at t = 0
    int done_it;         // PCG iterations >= done_it of the current GN step are skipped
    int pcg_iters;       // total PCG iterations executed
    int first;           // 1 until E0 has been recorded
};

DFU_DEV double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = l < (blockDim.x >> 5) ? sh[l] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;  // valid in thread 0
}

// fixed-order sum of per-block partials, computed redundantly by every block (deterministic)
DFU_DEV double sum_partials(const double* __restrict__ part, int n, double* sh) {
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += part[i];
    double r = block_sum(v, sh);
    __shared__ double bc;
    if (threadIdx.x == 0) bc = r;
    __syncthreads();
    return bc;
}

// calcTukeyBiweight (src/dynfu/utils/opt_solver.cpp:204-212)
DFU_DEV float tukey_biweight(float tukey_offset, float c, float ex, float ey, float ez) {
    const float s = __fdiv_rn(__fsqrt_rn(fadd(fadd(fmul(ex, ex), fmul(ey, ey)), fmul(ez, ez))), tukey_offset);
    if (s < c) {
        const double q = 1.0 - ((double) s * (double) s) / ((double) c * (double) c);
        return (float) (q * q);
    }
    return 0.f;
}

struct PointData {
    const int32_t* nbr;
    const float* wts;
    const float* dvec;
    float* theta;
    int P;
};

DFU_DEV void load8(const int32_t* nbr, const float* wts, int v, int (&nb)[8], float (&w)[8]) {
    const int4 a = __ldg(reinterpret_cast<const int4*>(nbr) + 2 * (size_t) v);
    const int4 b = __ldg(reinterpret_cast<const int4*>(nbr) + 2 * (size_t) v + 1);
    const float4 c = __ldg(reinterpret_cast<const float4*>(wts) + 2 * (size_t) v);
    const float4 d = __ldg(reinterpret_cast<const float4*>(wts) + 2 * (size_t) v + 1);
    nb[0] = a.x; nb[1] = a.y; nb[2] = a.z; nb[3] = a.w; nb[4] = b.x; nb[5] = b.y; nb[6] = b.z; nb[7] = b.w;
    w[0] = c.x; w[1] = c.y; w[2] = c.z; w[3] = c.w; w[4] = d.x; w[5] = d.y; w[6] = d.z; w[7] = d.w;
}

// Residual + Jacobian evaluation and per-node block assembly (data term).  J_vk = -sqrt(tukey) w_vk I3, so
// b[n] += tukey w e, D[n] += tukey w^2 (the 3x3 diagonal block is D*I3), E += tukey |e|^2.
__global__ void __launch_bounds__(TPB) assemble_data_kernel(PointData pd, const float* __restrict__ t, float* __restrict__ nbuf,
                                                            int N, int update_tukey, float tukey_offset, float psi_data) {
    __shared__ double sh[TPB / 32];
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    double e2 = 0.0;
    if (v < pd.P) {
        int nb[8];
        float w[8];
        load8(pd.nbr, pd.wts, v, nb, w);
        float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float* tk = t + 3 * (size_t) nb[k];
            sx = __fmaf_rn(w[k], tk[0], sx);
            sy = __fmaf_rn(w[k], tk[1], sy);
            sz = __fmaf_rn(w[k], tk[2], sz);
        }
        const float ex = pd.dvec[3 * (size_t) v] - sx, ey = pd.dvec[3 * (size_t) v + 1] - sy, ez = pd.dvec[3 * (size_t) v + 2] - sz;
        float th;
        if (update_tukey) {
            th = tukey_biweight(tukey_offset, psi_data, ex, ey, ez);
            pd.theta[v] = th;
        } else {
            th = pd.theta[v];
        }
        e2 = (double) th * ((double) ex * ex + (double) ey * ey + (double) ez * ez);
        if (th != 0.f) {
            float* b = nbuf;
            float* D = nbuf + 3 * (size_t) N;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float c = th * w[k];
                if (c != 0.f) {
                    atomicAdd(b + 3 * (size_t) nb[k], c * ex);
                    atomicAdd(b + 3 * (size_t) nb[k] + 1, c * ey);
                    atomicAdd(b + 3 * (size_t) nb[k] + 2, c * ez);
                    atomicAdd(D + nb[k], c * w[k]);
                }
            }
        }
    }
    const double bs = block_sum(e2, sh);
    if (threadIdx.x == 0 && bs != 0.0) atomicAdd(nbuf + 4 * (size_t) N, (float) bs);
}

struct RegGraph {
    const int32_t* nnbr;   // N*8 out-edges (n -> m)
    const int* rin_ptr;    // N+1
    const int32_t* rin;    // in-edges (sources m' that list n), sorted ascending
    float wreg2;
};

// regularisation part of b, D, E -- gather form over out- and in-edges, no atomics, fixed order
__global__ void __launch_bounds__(TPB) assemble_reg_kernel(RegGraph rg, const float* __restrict__ t, float* __restrict__ nbuf,
                                                           int N, double* __restrict__ partE) {
    __shared__ double sh[TPB / 32];
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (n < N && rg.wreg2 > 0.f) {
        const float tx = t[3 * (size_t) n], ty = t[3 * (size_t) n + 1], tz = t[3 * (size_t) n + 2];
        float bx = 0.f, by = 0.f, bz = 0.f, dd = 0.f;
        for (int i = 0; i < 8; ++i) {
            const int m = rg.nnbr[(size_t) n * 8 + i];
            if (m == n) continue;
            const float dx = t[3 * (size_t) m] - tx, dy = t[3 * (size_t) m + 1] - ty, dz = t[3 * (size_t) m + 2] - tz;
            bx += dx; by += dy; bz += dz;
            dd += 1.f;
            e += (double) dx * dx + (double) dy * dy + (double) dz * dz;
        }
        for (int j = rg.rin_ptr[n]; j < rg.rin_ptr[n + 1]; ++j) {
            const int m = rg.rin[j];
            if (m == n) continue;
            bx += t[3 * (size_t) m] - tx; by += t[3 * (size_t) m + 1] - ty; bz += t[3 * (size_t) m + 2] - tz;
            dd += 1.f;
        }
        nbuf[3 * (size_t) n] += rg.wreg2 * bx;
        nbuf[3 * (size_t) n + 1] += rg.wreg2 * by;
        nbuf[3 * (size_t) n + 2] += rg.wreg2 * bz;
        nbuf[3 * (size_t) N + n] += rg.wreg2 * dd;
        e *= (double) rg.wreg2;
    }
    const double bs = block_sum(e, sh);
    if (threadIdx.x == 0) partE[blockIdx.x] = bs;
}

struct Vecs {
    float *t, *dl, *r, *z, *p, *q;
    const float* nbuf;
    int N;
};

// r = b, z = M^-1 r, p = z, delta = 0, q = 0, partial r.z
__global__ void __launch_bounds__(TPB) pcg_init_kernel(Vecs x, double* __restrict__ part_rz) {
    __shared__ double sh[TPB / 32];
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    double rz = 0.0;
    if (n < x.N) {
        const float D = x.nbuf[3 * (size_t) x.N + n];
        const float inv = D > 0.f ? 1.f / D : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const size_t i = 3 * (size_t) n + c;
            const float r = x.nbuf[i], z = r * inv;
            x.r[i] = r; x.z[i] = z; x.p[i] = z; x.dl[i] = 0.f; x.q[i] = 0.f;
            rz += (double) r * z;
        }
    }
    const double bs = block_sum(rz, sh);
    if (threadIdx.x == 0) part_rz[blockIdx.x] = bs;
}

// single-thread bookkeeping after pcg_init: total energy, r.z, convergence of this GN step
__global__ void pcg_init_scalars_kernel(Scalars* sc, const float* __restrict__ nbuf, int N, const double* __restrict__ partE,
                                        const double* __restrict__ part_rz, int nblk, double tol2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double E = (double) nbuf[4 * (size_t) N], rz = 0.0;
    for (int i = 0; i < nblk; ++i) {
        E += partE[i];
        rz += part_rz[i];
    }
    sc->E = E;
    if (sc->first) {
        sc->E0 = E;
        sc->first = 0;
    }
    if (sc->rz_ref < 0.0) sc->rz_ref = rz;
    sc->rz[0] = rz;
    sc->done_it = (!(rz > 0.0) || rz <= tol2 * sc->rz_ref) ? 0 : INT_MAX;
}

// q += W^T Theta W p   (this rank's points)
__global__ void __launch_bounds__(TPB) apply_data_kernel(PointData pd, const float* __restrict__ p, float* __restrict__ q,
                                                         const Scalars* __restrict__ sc, int it) {
    if (it >= sc->done_it) return;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= pd.P) return;
    const float th = pd.theta[v];
    if (th == 0.f) return;
    int nb[8];
    float w[8];
    load8(pd.nbr, pd.wts, v, nb, w);
    float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float* pk = p + 3 * (size_t) nb[k];
        sx = __fmaf_rn(w[k], pk[0], sx);
        sy = __fmaf_rn(w[k], pk[1], sy);
        sz = __fmaf_rn(w[k], pk[2], sz);
    }
    sx *= th; sy *= th; sz *= th;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (w[k] != 0.f) {
            atomicAdd(q + 3 * (size_t) nb[k], w[k] * sx);
            atomicAdd(q + 3 * (size_t) nb[k] + 1, w[k] * sy);
            atomicAdd(q + 3 * (size_t) nb[k] + 2, w[k] * sz);
        }
    }
}

// q += w_reg^2 L p (gather, fixed order), partial p.q
__global__ void __launch_bounds__(TPB) apply_reg_dot_kernel(RegGraph rg, Vecs x, const Scalars* __restrict__ sc, int it,
                                                            double* __restrict__ part_pq) {
    if (it >= sc->done_it) return;
    __shared__ double sh[TPB / 32];
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    double pq = 0.0;
    if (n < x.N) {
        const float px = x.p[3 * (size_t) n], py = x.p[3 * (size_t) n + 1], pz = x.p[3 * (size_t) n + 2];
        float qx = x.q[3 * (size_t) n], qy = x.q[3 * (size_t) n + 1], qz = x.q[3 * (size_t) n + 2];
        if (rg.wreg2 > 0.f) {
            float ax = 0.f, ay = 0.f, az = 0.f;
            for (int i = 0; i < 8; ++i) {
                const int m = rg.nnbr[(size_t) n * 8 + i];
                if (m == n) continue;
                ax += px - x.p[3 * (size_t) m]; ay += py - x.p[3 * (size_t) m + 1]; az += pz - x.p[3 * (size_t) m + 2];
            }
            for (int j = rg.rin_ptr[n]; j < rg.rin_ptr[n + 1]; ++j) {
                const int m = rg.rin[j];
                if (m == n) continue;
                ax += px - x.p[3 * (size_t) m]; ay += py - x.p[3 * (size_t) m + 1]; az += pz - x.p[3 * (size_t) m + 2];
            }
            qx += rg.wreg2 * ax; qy += rg.wreg2 * ay; qz += rg.wreg2 * az;
            x.q[3 * (size_t) n] = qx; x.q[3 * (size_t) n + 1] = qy; x.q[3 * (size_t) n + 2] = qz;
        }
        pq = (double) px * qx + (double) py * qy + (double) pz * qz;
    }
    const double bs = block_sum(pq, sh);
    if (threadIdx.x == 0) part_pq[blockIdx.x] = bs;
}

// alpha = r.z / p.q ; delta += alpha p ; r -= alpha q ; z = M^-1 r ; q = 0 ; partial r.z
__global__ void __launch_bounds__(TPB) pcg_update_kernel(Vecs x, const Scalars* __restrict__ sc, int it,
                                                         const double* __restrict__ part_pq, int nblk,
                                                         double* __restrict__ part_rz) {
    if (it >= sc->done_it) return;
    __shared__ double sh[TPB / 32];
    const double pq = sum_partials(part_pq, nblk, sh);
    const double rz = sc->rz[it & 1];
    const float alpha = pq > 0.0 ? (float) (rz / pq) : 0.f;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    double rzn = 0.0;
    if (n < x.N) {
        const float D = x.nbuf[3 * (size_t) x.N + n];
        const float inv = D > 0.f ? 1.f / D : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const size_t i = 3 * (size_t) n + c;
            x.dl[i] = __fmaf_rn(alpha, x.p[i], x.dl[i]);
            const float r = __fmaf_rn(-alpha, x.q[i], x.r[i]);
            const float z = r * inv;
            x.r[i] = r; x.z[i] = z; x.q[i] = 0.f;
            rzn += (double) r * z;
        }
    }
    const double bs = block_sum(rzn, sh);
    if (threadIdx.x == 0) part_rz[blockIdx.x] = bs;
}

// beta = r.z_new / r.z ; p = z + beta p ; block 0 publishes r.z_new and the stop decision for it+1
__global__ void __launch_bounds__(TPB) pcg_direction_kernel(Vecs x, Scalars* sc, int it, const double* __restrict__ part_rz,
                                                            const double* __restrict__ part_pq, int nblk, double tol2) {
    if (it >= sc->done_it) return;
    __shared__ double sh[TPB / 32];
    const double rzn = sum_partials(part_rz, nblk, sh);
    const double rz = sc->rz[it & 1];
    const float beta = rz > 0.0 ? (float) (rzn / rz) : 0.f;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < x.N) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const size_t i = 3 * (size_t) n + c;
            x.p[i] = __fmaf_rn(beta, x.p[i], x.z[i]);
        }
    }
    if (blockIdx.x == 0) {
        const double pq = sum_partials(part_pq, nblk, sh);
        if (threadIdx.x == 0) {
            sc->rz[(it + 1) & 1] = rzn;
            sc->pcg_iters += 1;
            if (!(pq > 0.0) || !(rzn > 0.0) || rzn <= tol2 * sc->rz_ref) sc->done_it = it + 1;
        }
    }
}

__global__ void gn_update_kernel(float* __restrict__ t, const float* __restrict__ dl, int n3) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) t[i] += dl[i];
}

// ---- transposed regularisation graph (in-edges), deterministic ------------------------------------
__global__ void reg_indegree_kernel(const int32_t* __restrict__ nnbr, int N, int* __restrict__ deg) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < N * 8) atomicAdd(&deg[nnbr[e]], 1);
}
// exclusive scan of deg[0..N) -> ptr[0..N], single block
__global__ void __launch_bounds__(1024) scan_kernel(const int* __restrict__ deg, int N, int* __restrict__ ptr) {
    __shared__ int sh[1024];
    const int per = (N + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(N, lo + per);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += deg[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
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
__global__ void reg_fill_kernel(const int32_t* __restrict__ nnbr, int N, const int* __restrict__ ptr, int* __restrict__ cursor,
                                int32_t* __restrict__ rin) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < N * 8) {
        const int m = nnbr[e];
        rin[ptr[m] + atomicAdd(&cursor[m], 1)] = e / 8;
    }
}
__global__ void reg_sort_kernel(const int* __restrict__ ptr, int N, int32_t* __restrict__ rin) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int lo = ptr[n], hi = ptr[n + 1];
    for (int i = lo + 1; i < hi; ++i) {  // insertion sort: in-degree is small
        const int key = rin[i];
        int j = i - 1;
        while (j >= lo && rin[j] > key) {
            rin[j + 1] = rin[j];
            --j;
        }
        rin[j + 1] = key;
    }
}

}  // namespace

struct dfu_solver {
    dfu_warpfield* wf = nullptr;
    dfu_solver_params prm{};
    dfu_allreduce_fn allreduce = nullptr;
    void* allreduce_ctx = nullptr;
    int N = 0, P = 0;
    size_t capP = 0, capN = 0;
    // per point
    int32_t* nbr = nullptr;
    float *wts = nullptr, *dvec = nullptr, *theta = nullptr;
    // per node
    int32_t* nnbr = nullptr;
    int *rin_ptr = nullptr, *rin_tmp = nullptr;  // rin_tmp: [deg N | cursor N]
    int32_t* rin = nullptr;
    float* vec = nullptr;   // t, dl, r, z, p, q : 6 * 3N
    float* nbuf = nullptr;  // 4N + 4
    double* part = nullptr;  // 3 * MAX_PARTIALS : E, rz, pq
    Scalars* sc = nullptr;
    Scalars* sc_host = nullptr;  // pinned
    bool problem_ready = false;
    int gn_steps = 0;  // GN steps launched by the last solve_all
};

namespace {

int free_point_arrays(dfu_solver* s) {
    cudaFree(s->nbr); cudaFree(s->wts); cudaFree(s->dvec); cudaFree(s->theta);
    s->nbr = nullptr; s->wts = s->dvec = s->theta = nullptr;
    s->capP = 0;
    return DFU_OK;
}
int free_node_arrays(dfu_solver* s) {
    cudaFree(s->nnbr); cudaFree(s->rin_ptr); cudaFree(s->rin_tmp); cudaFree(s->rin); cudaFree(s->vec); cudaFree(s->nbuf);
    s->nnbr = nullptr; s->rin_ptr = s->rin_tmp = nullptr; s->rin = nullptr; s->vec = s->nbuf = nullptr;
    s->capN = 0;
    return DFU_OK;
}

int read_scalars(dfu_solver* s, cudaStream_t st) {
    DFU_CUDA_OK(cudaMemcpyAsync(s->sc_host, s->sc, sizeof(Scalars), cudaMemcpyDeviceToHost, st));
    DFU_CUDA_OK(cudaStreamSynchronize(st));
    return DFU_OK;
}

Vecs make_vecs(const dfu_solver* s) {
    const size_t n3 = 3 * (size_t) s->N;
    Vecs x;
    x.t = s->vec; x.dl = s->vec + n3; x.r = s->vec + 2 * n3; x.z = s->vec + 3 * n3; x.p = s->vec + 4 * n3; x.q = s->vec + 5 * n3;
    x.nbuf = s->nbuf;
    x.N = s->N;
    return x;
}

// one residual/Jacobian evaluation + block assembly (+ all-reduce) + regularisation + PCG initialisation
int assemble(dfu_solver* s, bool update_tukey, cudaStream_t st) {
    const int N = s->N, P = s->P, nblk = div_up(N, TPB);
    Vecs x = make_vecs(s);
    PointData pd{s->nbr, s->wts, s->dvec, s->theta, P};
    RegGraph rg{s->nnbr, s->rin_ptr, s->rin, s->prm.lambda / ((float) N * 8.f)};
    const size_t nbuf_count = 4 * (size_t) N + 4;
    DFU_CUDA_OK(cudaMemsetAsync(s->nbuf, 0, nbuf_count * sizeof(float), st));
    if (P > 0) {
        assemble_data_kernel<<<div_up(P, TPB), TPB, 0, st>>>(pd, x.t, s->nbuf, N, update_tukey ? 1 : 0, s->prm.tukey_offset,
                                                            s->prm.psi_data);
        DFU_LAUNCH_OK();
    }
    if (s->allreduce) {
        int rc = s->allreduce(s->nbuf, nbuf_count, s->allreduce_ctx, (dfu_stream) st);
        DFU_REQUIRE(rc == 0, DFU_ERR_CUDA, "all-reduce hook failed");
    }
    assemble_reg_kernel<<<nblk, TPB, 0, st>>>(rg, x.t, s->nbuf, N, s->part);
    DFU_LAUNCH_OK();
    pcg_init_kernel<<<nblk, TPB, 0, st>>>(x, s->part + MAX_PARTIALS);
    DFU_LAUNCH_OK();
    const double tol2 = (double) s->prm.pcg_tol * (double) s->prm.pcg_tol;
    pcg_init_scalars_kernel<<<1, 32, 0, st>>>(s->sc, s->nbuf, N, s->part, s->part + MAX_PARTIALS, nblk, tol2);
    DFU_LAUNCH_OK();
    return DFU_OK;
}

}  // namespace

extern "C" {

int dfu_solver_create(dfu_solver** out, dfu_warpfield* wf, const dfu_solver_params* prm) {
    DFU_REQUIRE(out && wf && prm, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(prm->num_iter >= 0 && prm->nonlinear_iter >= 0 && prm->linear_iter >= 0, DFU_ERR_INVALID, "negative iteration count");
    DFU_REQUIRE(prm->tukey_offset > 0.f && prm->psi_data > 0.f && prm->lambda >= 0.f, DFU_ERR_INVALID, "bad robust/regularisation parameter");
    dfu_solver* s = new dfu_solver();
    s->wf = wf;
    s->prm = *prm;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(wf->device);
    cudaError_t e1 = cudaMalloc(&s->sc, sizeof(Scalars));
    cudaError_t e2 = cudaMallocHost(&s->sc_host, sizeof(Scalars));
    cudaError_t e3 = cudaMalloc(&s->part, 3 * MAX_PARTIALS * sizeof(double));
    cudaSetDevice(prev);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        dfu_set_error("dfu_solver_create: allocation failed");
        delete s;
        return DFU_ERR_CUDA;
    }
    *out = s;
    return DFU_OK;
}

int dfu_solver_destroy(dfu_solver* s) {
    if (!s) return DFU_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(s->wf->device);
    free_point_arrays(s);
    free_node_arrays(s);
    cudaFree(s->sc);
    cudaFreeHost(s->sc_host);
    cudaFree(s->part);
    cudaSetDevice(prev);
    delete s;
    return DFU_OK;
}

int dfu_solver_set_allreduce(dfu_solver* s, dfu_allreduce_fn fn, void* ctx) {
    DFU_REQUIRE(s, DFU_ERR_INVALID, "NULL argument");
    s->allreduce = fn;
    s->allreduce_ctx = ctx;
    return DFU_OK;
}

int dfu_solver_init_problem(dfu_solver* s, const float* canon_v, const float* canon_n, const float* live_v,
                            const float* live_n, int P, const float affine_host[12], dfu_stream stream) {
    (void) canon_n; (void) live_n; (void) affine_host;  // uploaded but never read by the reference's energy
    DFU_REQUIRE(s, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(P >= 0 && (P == 0 || (canon_v && live_v)), DFU_ERR_INVALID, "bad point arrays");
    dfu_warpfield* wf = s->wf;
    DFU_REQUIRE(wf->initialised, DFU_ERR_NOT_INIT, "warp field not initialised");
    DFU_REQUIRE(wf->N >= DFU_KNN, DFU_ERR_PRECONDITION, "the solver needs at least 8 nodes (reference UB, opt_solver.cpp:63-66)");
    DFU_REQUIRE(div_up(wf->N, TPB) <= MAX_PARTIALS, DFU_ERR_INVALID, "too many nodes");
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != wf->device) DFU_CUDA_OK(cudaSetDevice(wf->device));
    cudaStream_t st = as_stream(stream);
    const int N = wf->N;
    if ((size_t) P > s->capP) {
        free_point_arrays(s);
        const size_t cap = (size_t) P;
        DFU_CUDA_OK(cudaMalloc(&s->nbr, cap * 8 * sizeof(int32_t)));
        DFU_CUDA_OK(cudaMalloc(&s->wts, cap * 8 * sizeof(float)));
        DFU_CUDA_OK(cudaMalloc(&s->dvec, cap * 3 * sizeof(float)));
        DFU_CUDA_OK(cudaMalloc(&s->theta, cap * sizeof(float)));
        s->capP = cap;
    }
    if ((size_t) N > s->capN) {
        free_node_arrays(s);
        const size_t cap = (size_t) N;
        DFU_CUDA_OK(cudaMalloc(&s->nnbr, cap * 8 * sizeof(int32_t)));
        DFU_CUDA_OK(cudaMalloc(&s->rin_ptr, (cap + 1) * sizeof(int)));
        DFU_CUDA_OK(cudaMalloc(&s->rin_tmp, 2 * cap * sizeof(int)));
        DFU_CUDA_OK(cudaMalloc(&s->rin, cap * 8 * sizeof(int32_t)));
        DFU_CUDA_OK(cudaMalloc(&s->vec, 18 * cap * sizeof(float)));
        DFU_CUDA_OK(cudaMalloc(&s->nbuf, (4 * cap + 4) * sizeof(float)));
        s->capN = cap;
    }
    s->N = N;
    s->P = P;
    int rc = DFU_OK;
    if (P > 0) rc = dfu_wf_build_data_graph(wf, canon_v, live_v, P, s->nbr, s->wts, s->dvec, st);  // opt_solver.cpp:56-72
    if (rc == DFU_OK) rc = dfu_wf_build_node_graph(wf, s->nnbr, st);                             // opt_solver.cpp:74-105
    if (rc != DFU_OK) return rc;
    DFU_CUDA_OK(cudaMemsetAsync(s->rin_tmp, 0, 2 * (size_t) N * sizeof(int), st));
    reg_indegree_kernel<<<div_up((long) N * 8, TPB), TPB, 0, st>>>(s->nnbr, N, s->rin_tmp);
    DFU_LAUNCH_OK();
    scan_kernel<<<1, 1024, 0, st>>>(s->rin_tmp, N, s->rin_ptr);
    DFU_LAUNCH_OK();
    reg_fill_kernel<<<div_up((long) N * 8, TPB), TPB, 0, st>>>(s->nnbr, N, s->rin_ptr, s->rin_tmp + N, s->rin);
    DFU_LAUNCH_OK();
    reg_sort_kernel<<<div_up(N, TPB), TPB, 0, st>>>(s->rin_ptr, N, s->rin);
    DFU_LAUNCH_OK();
    DFU_CUDA_OK(cudaMemsetAsync(s->vec, 0, 18 * (size_t) N * sizeof(float), st));  // unknowns := 0 (opt_solver.cpp:192-193)
    s->problem_ready = true;
    if (prev != wf->device) cudaSetDevice(prev);
    return DFU_OK;
}

int dfu_solver_solve_all(dfu_solver* s, dfu_stream stream) {
    DFU_REQUIRE(s, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(s->problem_ready, DFU_ERR_NOT_INIT, "initializeProblemInstance has not been called");
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != s->wf->device) DFU_CUDA_OK(cudaSetDevice(s->wf->device));
    cudaStream_t st = as_stream(stream);
    const int N = s->N, P = s->P, nblk = div_up(N, TPB);
    const dfu_solver_params& prm = s->prm;
    Vecs x = make_vecs(s);
    PointData pd{s->nbr, s->wts, s->dvec, s->theta, P};
    RegGraph rg{s->nnbr, s->rin_ptr, s->rin, prm.lambda / ((float) N * 8.f)};
    const double tol2 = (double) prm.pcg_tol * (double) prm.pcg_tol;
    const bool host_checks = prm.early_out != 0;
    double* partE = s->part;
    double* part_rz = s->part + MAX_PARTIALS;
    double* part_pq = s->part + 2 * MAX_PARTIALS;

    Scalars init{};
    init.rz_ref = -1.0;
    init.first = 1;
    init.done_it = INT_MAX;
    *s->sc_host = init;
    DFU_CUDA_OK(cudaMemcpyAsync(s->sc, s->sc_host, sizeof(Scalars), cudaMemcpyHostToDevice, st));
    DFU_CUDA_OK(cudaMemsetAsync(x.t, 0, 3 * (size_t) N * sizeof(float), st));
    if (host_checks) DFU_CUDA_OK(cudaStreamSynchronize(st));  // sc_host is reused for read-backs below

    bool stop_all = false;
    s->gn_steps = 0;
    for (int outer = 0; outer < prm.num_iter && !stop_all; ++outer) {
        for (int gn = 0; gn < prm.nonlinear_iter; ++gn) {
            // preNonlinearSolve re-weights once per outer iteration (opt_solver.cpp:135-140)
            int rc = assemble(s, gn == 0, st);
            if (rc != DFU_OK) return rc;
            if (host_checks) {
                rc = read_scalars(s, st);
                if (rc != DFU_OK) return rc;
                if (s->sc_host->done_it == 0) {  // already converged at this linearisation point
                    if (gn == 0 && outer > 0) stop_all = true;
                    break;
                }
            }
            for (int it = 0; it < prm.linear_iter; ++it) {
                if (P > 0) {
                    apply_data_kernel<<<div_up(P, TPB), TPB, 0, st>>>(pd, x.p, x.q, s->sc, it);
                    DFU_LAUNCH_OK();
                }
                if (s->allreduce) {
                    rc = s->allreduce(x.q, 3 * (size_t) N, s->allreduce_ctx, (dfu_stream) st);
                    DFU_REQUIRE(rc == 0, DFU_ERR_CUDA, "all-reduce hook failed");
                }
                apply_reg_dot_kernel<<<nblk, TPB, 0, st>>>(rg, x, s->sc, it, part_pq);
                DFU_LAUNCH_OK();
                pcg_update_kernel<<<nblk, TPB, 0, st>>>(x, s->sc, it, part_pq, nblk, part_rz);
                DFU_LAUNCH_OK();
                pcg_direction_kernel<<<nblk, TPB, 0, st>>>(x, s->sc, it, part_rz, part_pq, nblk, tol2);
                DFU_LAUNCH_OK();
                if (host_checks && (it & 7) == 7) {
                    rc = read_scalars(s, st);
                    if (rc != DFU_OK) return rc;
                    if (s->sc_host->done_it <= it + 1) break;
                }
            }
            gn_update_kernel<<<div_up(3L * N, TPB), TPB, 0, st>>>(x.t, x.dl, 3 * N);
            DFU_LAUNCH_OK();
            s->gn_steps += 1;
        }
    }
    // final energy at the solution (Tukey weights of the last outer iteration), then write back ONCE:
    // dg_se3 := DQ(0,0,0,t) * dg_se3 (opt_solver.cpp:270-285, node.cpp:19-23)
    int rc = assemble(s, false, st);
    if (rc != DFU_OK) return rc;
    rc = dfu_warpfield_update_translations(s->wf, x.t, stream);
    if (prev != s->wf->device) cudaSetDevice(prev);
    (void) partE;
    return rc;
}

int dfu_solver_get_translations(const dfu_solver* s, float* t_xyz, dfu_stream stream) {
    DFU_REQUIRE(s && t_xyz, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(s->problem_ready, DFU_ERR_NOT_INIT, "no problem instance");
    DFU_CUDA_OK(cudaMemcpyAsync(t_xyz, s->vec, 3 * (size_t) s->N * sizeof(float), cudaMemcpyDeviceToDevice, as_stream(stream)));
    return DFU_OK;
}

int dfu_solver_get_stats_host(const dfu_solver* s, double stats_host[4], dfu_stream stream) {
    DFU_REQUIRE(s && stats_host, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(s->problem_ready, DFU_ERR_NOT_INIT, "no problem instance");
    int rc = read_scalars(const_cast<dfu_solver*>(s), as_stream(stream));
    if (rc != DFU_OK) return rc;
    stats_host[0] = s->sc_host->E0;
    stats_host[1] = s->sc_host->E;
    stats_host[2] = (double) s->sc_host->pcg_iters;
    stats_host[3] = (double) s->gn_steps;
    return DFU_OK;
}

}  // extern "C"
