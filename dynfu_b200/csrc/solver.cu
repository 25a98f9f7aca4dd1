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
// Data layout in HBM (all L2-resident at the sizes of BASELINE.json):
//   per point   nbr[8] i32, wts[8] f32 (two 128-bit loads each), dvec = live - canon, theta (tukey),
//               s4 = float4 scratch (theta*e | theta  during assembly, theta*W p during PCG)
//   transposed  tptr[N+1], tv[8P], tw[8P]: for every node the (point, weight) pairs that reference it,
//               sorted by point -- the per-node J^T J / J^T r blocks and the PCG product W^T(...) are GATHERS
//               over these lists in a fixed order: no atomics in the iteration loop, bit-reproducible
//   per node    t, delta, r, z, p, q (3 floats), nbuf = [ b = -J^T r : 3N | D = diag(J^T J) : N | E : 4 ]
//
// Two execution paths, same arithmetic phases:
//   * single rank: ONE persistent cooperative kernel runs every outer / GN / PCG iteration with grid-wide
//     barriers between phases (3 per PCG iteration) and decides convergence on the device -- no launches and
//     no host round trips inside solveAll;
//   * data-parallel ranks (all-reduce hook set): one kernel per phase, nbuf all-reduced once per GN step
//     and q (3N floats) once per PCG iteration; everything after an all-reduce is evaluated in a fixed
//     order, so all ranks compute bit-identical iterates.
#include <cooperative_groups.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "dfu_internal.h"
#include "dfu_math.cuh"

namespace cg = cooperative_groups;
using namespace dfu;

namespace {

constexpr int TPB = 256;           // multi-kernel path
constexpr int PTPB = 512;          // persistent kernel
constexpr int MAX_PARTIALS = 1024;

struct Scalars {
    double rz[2];   // multi-kernel path: r.z of PCG iteration it is rz[it & 1]
    double rz_ref;  // r.z of the first GN step of this solve (< 0: unset)
    double E;       // energy at the last evaluation (data + reg)
    double E0;      // energy at t = 0
    int done_it;    // multi-kernel path: PCG iterations >= done_it of the current GN step are skipped
    int pcg_iters;  // total PCG iterations executed
    int gn_steps;   // total GN steps executed
    int first;      // 1 until E0 has been recorded
    int spin_fail;  // version 3r: a tagged word never arrived (would have been a hang); the result is invalid
};

struct Problem {
    int N, P;
    // data graph
    const int32_t* nbr;
    const float* wts;
    const float* dvec;
    float* theta;
    float4* s4;
    const int* tptr;
    const int32_t* tv;
    const float* tw;
    // regularisation graph: out-edges n -> nnbr[n][i], in-edges rin[rin_ptr[n]..) (sources, ascending)
    const int32_t* nnbr;
    const int* rin_ptr;
    const int32_t* rin;
    float wreg2;
    // unknowns and PCG vectors
    float *t, *dl, *r, *z, *p, *q;
    float* nbuf;   // b [3N] | D [N] | E [4]
    double* part;  // 4 * MAX_PARTIALS
    float tukey_offset, psi_data;
};

// Explicit normal matrix of one GN step, A = W^T Theta W + w_reg^2 L (N x N, the same for the 3 coordinates): CSR-like
// rows stored in arbitrary order (rowptr/rowlen), columns ascending inside a row.  The sparsity pattern and the
// regularisation part are built once per frame (k_pattern), the data part once per re-weighting.
struct Pattern {
    const int* rowptr;
    const int* rowlen;
    const int* dslot;       // slot of the diagonal entry in its row
    const int32_t* col;
    const float* areg;      // w_reg^2 * L
    float* vals;            // areg + W^T Theta W
    const uint4* tslot;     // per transposed-graph entry (node a, point v): slots in row a of v's 8 neighbours (8 x u16)
    float4* exch;           // [2][N] vector exchanged between CTAs (u0 / m_i), double buffered
    float4* st;             // [6][N] row-local PCG state: r, w, z, s, p, x
    unsigned long long* xw; // register version: [2][N][3] exchanged vector as tagged words (float bits | tag << 32)
    unsigned long long* pw; // register version: [2][MAX_PARTIALS][2] tagged per-CTA partial sums
};
constexpr int ACC_W = 256;                              // fixed-point accumulators per warp (columns per pass)
constexpr float FIX_SCALE = 1099511627776.f;            // 2^40; contributions are <= 1
constexpr double FIX_INV = 1.0 / 1099511627776.0;
// the register version splits a contribution into a 20-bit low and a 21-bit high word and adds them with native 32-bit
// shared-memory atomics: exact (integer) as long as a row collects fewer than 2048 contributions per column
constexpr int FIX_MAX_DEG = 2047;
DFU_DEV float fix2f(unsigned lo, unsigned hi) { return (float) ((double) (((unsigned long long) hi << 20) + lo) * FIX_INV); }

// ---------------------------------------------------------------------------------------------------
DFU_DEV double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
DFU_DEV float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// sum over the CTA, result broadcast to every thread; fixed reduction tree (deterministic)
DFU_DEV double block_sum(double v, double* sh) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = l < (int) (blockDim.x >> 5) ? sh[l] : 0.0;
    return warp_sum(r);
}
// four sums at once (one pair of CTA barriers instead of four)
struct D4 {
    double a, b, c, d;
};
DFU_DEV D4 block_sum4(D4 v, double* sh4) {  // sh4: 4 * (blockDim/32) doubles
    v.a = warp_sum(v.a); v.b = warp_sum(v.b); v.c = warp_sum(v.c); v.d = warp_sum(v.d);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nwp = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) {
        sh4[w] = v.a; sh4[nwp + w] = v.b; sh4[2 * nwp + w] = v.c; sh4[3 * nwp + w] = v.d;
    }
    __syncthreads();
    D4 r;
    r.a = warp_sum(l < nwp ? sh4[l] : 0.0);
    r.b = warp_sum(l < nwp ? sh4[nwp + l] : 0.0);
    r.c = warp_sum(l < nwp ? sh4[2 * nwp + l] : 0.0);
    r.d = warp_sum(l < nwp ? sh4[3 * nwp + l] : 0.0);
    return r;
}
// per-CTA partials are stored as 4 consecutive doubles per CTA; every CTA sums them redundantly in a fixed
// order with ONE pass over memory
DFU_DEV D4 sum_partials4(const double* part4, int n, double* sh4) {
    D4 v{0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double2 x = *reinterpret_cast<const double2*>(part4 + 4 * (size_t) i);
        const double2 y = *reinterpret_cast<const double2*>(part4 + 4 * (size_t) i + 2);
        v.a += x.x; v.b += x.y; v.c += y.x; v.d += y.y;
    }
    return block_sum4(v, sh4);
}

// Grid-wide barrier of the persistent kernel: one release-add per CTA on a monotonically increasing counter
// (zeroed by the host before the launch) and an acquire-poll until all CTAs of this generation have arrived.
// The kernel is launched cooperatively, so all CTAs are co-resident.
DFU_DEV void grid_barrier(unsigned* counter, unsigned nblocks, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += nblocks;
        unsigned seen;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
    }
    __syncthreads();
}

// fixed-order sum of per-block partials, computed redundantly by every block
DFU_DEV double sum_partials(const double* part, int n, double* sh) {
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += part[i];
    return block_sum(v, sh);
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

DFU_DEV void load8(const int32_t* nbr, const float* wts, int v, int (&nb)[8], float (&w)[8]) {
    const int4 a = *(reinterpret_cast<const int4*>(nbr) + 2 * (size_t) v);
    const int4 b = *(reinterpret_cast<const int4*>(nbr) + 2 * (size_t) v + 1);
    const float4 c = *(reinterpret_cast<const float4*>(wts) + 2 * (size_t) v);
    const float4 d = *(reinterpret_cast<const float4*>(wts) + 2 * (size_t) v + 1);
    nb[0] = a.x; nb[1] = a.y; nb[2] = a.z; nb[3] = a.w; nb[4] = b.x; nb[5] = b.y; nb[6] = b.z; nb[7] = b.w;
    w[0] = c.x; w[1] = c.y; w[2] = c.z; w[3] = c.w; w[4] = d.x; w[5] = d.y; w[6] = d.z; w[7] = d.w;
}

// sum_k w_k x[n_k] for one point
DFU_DEV void point_gather(const Problem& pb, int v, const float* x, float& sx, float& sy, float& sz) {
    int nb[8];
    float w[8];
    load8(pb.nbr, pb.wts, v, nb, w);
    sx = sy = sz = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float* xk = x + 3 * (size_t) nb[k];
        sx = __fmaf_rn(w[k], xk[0], sx);
        sy = __fmaf_rn(w[k], xk[1], sy);
        sz = __fmaf_rn(w[k], xk[2], sz);
    }
}

// ---- phases (grid-stride; tid/nthreads describe the whole launch) --------------------------------------
// Residual evaluation (energy.t:47-55): e = d - W t, tukey re-weighting, s4 = (theta e, theta).
// Returns this thread's share of sum theta |e|^2.
DFU_DEV double phase_point_residual(const Problem& pb, bool update_tukey, int tid, int nthreads) {
    double e2 = 0.0;
    for (int v = tid; v < pb.P; v += nthreads) {
        float sx, sy, sz;
        point_gather(pb, v, pb.t, sx, sy, sz);
        const float ex = pb.dvec[3 * (size_t) v] - sx, ey = pb.dvec[3 * (size_t) v + 1] - sy,
                    ez = pb.dvec[3 * (size_t) v + 2] - sz;
        float th;
        if (update_tukey) {
            th = tukey_biweight(pb.tukey_offset, pb.psi_data, ex, ey, ez);
            pb.theta[v] = th;
        } else {
            th = pb.theta[v];
        }
        pb.s4[v] = make_float4(th * ex, th * ey, th * ez, th);
        e2 += (double) th * ((double) ex * ex + (double) ey * ey + (double) ez * ez);
    }
    return e2;
}

// s4 = theta * W p
DFU_DEV void phase_point_apply(const Problem& pb, int tid, int nthreads) {
    for (int v = tid; v < pb.P; v += nthreads) {
        const float th = pb.theta[v];
        float sx = 0.f, sy = 0.f, sz = 0.f;
        if (th != 0.f) point_gather(pb, v, pb.p, sx, sy, sz);
        pb.s4[v] = make_float4(th * sx, th * sy, th * sz, th);
    }
}

// per-node gather of the data term over the transposed graph (one warp per node): returns, in every lane,
// sum_j tw_j * s4[tv_j].xyz and (with_diag) sum_j tw_j^2 * s4[tv_j].w -- lane-strided, then a fixed xor tree
DFU_DEV void node_gather_data(const Problem& pb, int n, int lane, bool with_diag, float& ax, float& ay, float& az, float& ad) {
    ax = ay = az = ad = 0.f;
    const int lo = pb.tptr[n], hi = pb.tptr[n + 1];
    // 4 entries per lane in flight: all index/weight loads first, then the dependent s4 gathers, then the
    // accumulation in entry order (same order as a plain lane-strided loop -> same bits)
    for (int j = lo + lane; j < hi; j += 128) {
        float w[4];
        int v[4];
        float4 s[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int jj = j + 32 * u;
            const bool ok = jj < hi;
            w[u] = ok ? pb.tw[jj] : 0.f;
            v[u] = ok ? pb.tv[jj] : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) s[u] = v[u] >= 0 ? pb.s4[v[u]] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (v[u] < 0) continue;
            ax = __fmaf_rn(w[u], s[u].x, ax);
            ay = __fmaf_rn(w[u], s[u].y, ay);
            az = __fmaf_rn(w[u], s[u].z, az);
            if (with_diag) ad = __fmaf_rn(w[u] * w[u], s[u].w, ad);
        }
    }
}

// The same gather (without the diagonal) in 2^40 fixed point: every term is rounded to an integer and integers add
// associatively, so the result does not depend on the ORDER of the transposed list -- the lists then need no sorting
// (versions 3 / 3r).  Terms are |tw * theta * e| << 2^23, a node collects < 2^19 of them.  Returns the sums in every lane.
DFU_DEV void warp_sum_ll(long long& v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
}
DFU_DEV void node_gather_data_fixed(const Problem& pb, int n, int lane, float& ax, float& ay, float& az) {
    long long sx = 0, sy = 0, sz = 0;
    const int lo = pb.tptr[n], hi = pb.tptr[n + 1];
    for (int j = lo + lane; j < hi; j += 128) {
        float w[4];
        int v[4];
        float4 s[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int jj = j + 32 * u;
            const bool ok = jj < hi;
            w[u] = ok ? pb.tw[jj] : 0.f;
            v[u] = ok ? pb.tv[jj] : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) s[u] = v[u] >= 0 ? pb.s4[v[u]] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (v[u] < 0) continue;
            sx += __float2ll_rn(w[u] * s[u].x * FIX_SCALE);
            sy += __float2ll_rn(w[u] * s[u].y * FIX_SCALE);
            sz += __float2ll_rn(w[u] * s[u].z * FIX_SCALE);
        }
    }
    warp_sum_ll(sx); warp_sum_ll(sy); warp_sum_ll(sz);
    ax = (float) ((double) sx * FIX_INV); ay = (float) ((double) sy * FIX_INV); az = (float) ((double) sz * FIX_INV);
}

// lane-parallel regularisation gather for node n on vector x: sum over out- and in-edges (m != n) of
// (x[n] - x[m]) in (gx,gy,gz), the edge count in cnt and (out-edges only) the squared differences in e2
DFU_DEV void node_gather_reg(const Problem& pb, int n, int lane, const float* x, float& gx, float& gy, float& gz, float& cnt,
                             float& e2) {
    gx = gy = gz = cnt = e2 = 0.f;
    const float xn0 = x[3 * (size_t) n], xn1 = x[3 * (size_t) n + 1], xn2 = x[3 * (size_t) n + 2];
    const int lo = pb.rin_ptr[n], hi = pb.rin_ptr[n + 1];
    for (int j = lane; j < 8 + (hi - lo); j += 32) {
        const bool out = j < 8;
        const int m = out ? pb.nnbr[(size_t) n * 8 + j] : pb.rin[lo + j - 8];
        if (m == n) continue;
        const float d0 = xn0 - x[3 * (size_t) m], d1 = xn1 - x[3 * (size_t) m + 1], d2 = xn2 - x[3 * (size_t) m + 2];
        gx += d0; gy += d1; gz += d2;
        cnt += 1.f;
        if (out) e2 += d0 * d0 + d1 * d1 + d2 * d2;
    }
}

// =====================================================================================================
// multi-kernel path (data-parallel ranks with an all-reduce hook)

__global__ void __launch_bounds__(TPB) k_point_residual(Problem pb, int update_tukey) {
    __shared__ double sh[TPB / 32];
    const double e2 = phase_point_residual(pb, update_tukey != 0, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
    const double bs = block_sum(e2, sh);
    if (threadIdx.x == 0) pb.part[blockIdx.x] = bs;
}
// data part of b, D into nbuf; block 0 also folds the energy partials into nbuf[4N]
__global__ void __launch_bounds__(TPB) k_node_assemble_data(Problem pb, int n_epart) {
    __shared__ double sh[TPB / 32];
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int n = gw; n < pb.N; n += nw) {
        float ax, ay, az, ad;
        node_gather_data(pb, n, lane, true, ax, ay, az, ad);
        ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az); ad = warp_sum(ad);
        if (lane == 0) {
            pb.nbuf[3 * (size_t) n] = ax; pb.nbuf[3 * (size_t) n + 1] = ay; pb.nbuf[3 * (size_t) n + 2] = az;
            pb.nbuf[3 * (size_t) pb.N + n] = ad;
        }
    }
    if (blockIdx.x == 0) {
        const double E = sum_partials(pb.part, n_epart, sh);
        if (threadIdx.x == 0) {
            pb.nbuf[4 * (size_t) pb.N] = (float) E;
            pb.nbuf[4 * (size_t) pb.N + 1] = pb.nbuf[4 * (size_t) pb.N + 2] = pb.nbuf[4 * (size_t) pb.N + 3] = 0.f;
        }
    }
}
// after the all-reduce: regularisation part of b, D, E; r = b, z = M^-1 r, p = z, delta = 0; partial r.z, E_reg
__global__ void __launch_bounds__(TPB) k_node_reg_init(Problem pb) {
    __shared__ double sh[TPB / 32];
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    double rz = 0.0, er = 0.0;
    for (int n = gw; n < pb.N; n += nw) {
        float gx = 0.f, gy = 0.f, gz = 0.f, cnt = 0.f, e2 = 0.f;
        if (pb.wreg2 > 0.f) {
            node_gather_reg(pb, n, lane, pb.t, gx, gy, gz, cnt, e2);
            gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz); cnt = warp_sum(cnt); e2 = warp_sum(e2);
        }
        if (lane == 0) {
            // d/dt_n of w^2 |t_m - t_n|^2 (both edge directions): b gets w^2 * sum (t_m - t_n) = -w^2 * g
            const float b0 = pb.nbuf[3 * (size_t) n] - pb.wreg2 * gx, b1 = pb.nbuf[3 * (size_t) n + 1] - pb.wreg2 * gy,
                        b2 = pb.nbuf[3 * (size_t) n + 2] - pb.wreg2 * gz;
            const float D = pb.nbuf[3 * (size_t) pb.N + n] + pb.wreg2 * cnt;
            pb.nbuf[3 * (size_t) n] = b0; pb.nbuf[3 * (size_t) n + 1] = b1; pb.nbuf[3 * (size_t) n + 2] = b2;
            pb.nbuf[3 * (size_t) pb.N + n] = D;
            const float inv = D > 0.f ? 1.f / D : 0.f;
            const float bb[3] = {b0, b1, b2};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const size_t i = 3 * (size_t) n + c;
                const float z = bb[c] * inv;
                pb.r[i] = bb[c]; pb.z[i] = z; pb.p[i] = z; pb.dl[i] = 0.f;
                rz += (double) bb[c] * z;
            }
            er += (double) pb.wreg2 * e2;
        }
    }
    const double a = block_sum(rz, sh), b = block_sum(er, sh);
    if (threadIdx.x == 0) {
        pb.part[MAX_PARTIALS + blockIdx.x] = a;
        pb.part[2 * MAX_PARTIALS + blockIdx.x] = b;
    }
}
__global__ void k_init_scalars(Problem pb, Scalars* sc, int nblk, double tol2) {
    __shared__ double sh[1];
    if (blockIdx.x != 0) return;
    const double rz = sum_partials(pb.part + MAX_PARTIALS, nblk, sh);
    const double er = sum_partials(pb.part + 2 * MAX_PARTIALS, nblk, sh);
    if (threadIdx.x == 0) {
        const double E = (double) pb.nbuf[4 * (size_t) pb.N] + er;
        sc->E = E;
        if (sc->first) {
            sc->E0 = E;
            sc->first = 0;
        }
        if (sc->rz_ref < 0.0) sc->rz_ref = rz;
        sc->rz[0] = rz;
        sc->done_it = (!(rz > 0.0) || rz <= tol2 * sc->rz_ref) ? 0 : INT_MAX;
    }
}
__global__ void __launch_bounds__(TPB) k_point_apply(Problem pb, const Scalars* sc, int it) {
    if (it >= sc->done_it) return;
    phase_point_apply(pb, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
__global__ void __launch_bounds__(TPB) k_node_apply_data(Problem pb, const Scalars* sc, int it) {
    if (it >= sc->done_it) return;
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int n = gw; n < pb.N; n += nw) {
        float ax, ay, az, ad;
        node_gather_data(pb, n, lane, false, ax, ay, az, ad);
        ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
        if (lane == 0) {
            pb.q[3 * (size_t) n] = ax; pb.q[3 * (size_t) n + 1] = ay; pb.q[3 * (size_t) n + 2] = az;
        }
    }
}
// after the all-reduce of q: q += w_reg^2 L p, partial p.q
__global__ void __launch_bounds__(TPB) k_node_apply_reg_dot(Problem pb, const Scalars* sc, int it) {
    if (it >= sc->done_it) return;
    __shared__ double sh[TPB / 32];
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    double pq = 0.0;
    for (int n = gw; n < pb.N; n += nw) {
        float gx = 0.f, gy = 0.f, gz = 0.f, cnt = 0.f, e2 = 0.f;
        if (pb.wreg2 > 0.f) {
            node_gather_reg(pb, n, lane, pb.p, gx, gy, gz, cnt, e2);
            gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
        }
        if (lane == 0) {
            const float q0 = pb.q[3 * (size_t) n] + pb.wreg2 * gx, q1 = pb.q[3 * (size_t) n + 1] + pb.wreg2 * gy,
                        q2 = pb.q[3 * (size_t) n + 2] + pb.wreg2 * gz;
            pb.q[3 * (size_t) n] = q0; pb.q[3 * (size_t) n + 1] = q1; pb.q[3 * (size_t) n + 2] = q2;
            pq += (double) pb.p[3 * (size_t) n] * q0 + (double) pb.p[3 * (size_t) n + 1] * q1 + (double) pb.p[3 * (size_t) n + 2] * q2;
        }
    }
    const double bs = block_sum(pq, sh);
    if (threadIdx.x == 0) pb.part[blockIdx.x] = bs;
}
// alpha = r.z / p.q ; delta += alpha p ; r -= alpha q ; z = M^-1 r ; partial r.z
__global__ void __launch_bounds__(TPB) k_pcg_update(Problem pb, const Scalars* sc, int it, int nblk_pq) {
    if (it >= sc->done_it) return;
    __shared__ double sh[TPB / 32];
    const double pq = sum_partials(pb.part, nblk_pq, sh);
    const double rz = sc->rz[it & 1];
    const float alpha = pq > 0.0 ? (float) (rz / pq) : 0.f;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    double rzn = 0.0;
    if (n < pb.N) {
        const float D = pb.nbuf[3 * (size_t) pb.N + n];
        const float inv = D > 0.f ? 1.f / D : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const size_t i = 3 * (size_t) n + c;
            pb.dl[i] = __fmaf_rn(alpha, pb.p[i], pb.dl[i]);
            const float r = __fmaf_rn(-alpha, pb.q[i], pb.r[i]);
            const float z = r * inv;
            pb.r[i] = r; pb.z[i] = z;
            rzn += (double) r * z;
        }
    }
    const double bs = block_sum(rzn, sh);
    if (threadIdx.x == 0) pb.part[MAX_PARTIALS + blockIdx.x] = bs;
}
// beta = r.z_new / r.z ; p = z + beta p ; block 0 publishes r.z_new and the stop decision for it+1
__global__ void __launch_bounds__(TPB) k_pcg_direction(Problem pb, Scalars* sc, int it, int nblk, int nblk_pq, double tol2) {
    if (it >= sc->done_it) return;
    __shared__ double sh[TPB / 32];
    const double rzn = sum_partials(pb.part + MAX_PARTIALS, nblk, sh);
    const double rz = sc->rz[it & 1];
    const float beta = rz > 0.0 ? (float) (rzn / rz) : 0.f;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < pb.N) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const size_t i = 3 * (size_t) n + c;
            pb.p[i] = __fmaf_rn(beta, pb.p[i], pb.z[i]);
        }
    }
    if (blockIdx.x == 0) {
        const double pq = sum_partials(pb.part, nblk_pq, sh);
        if (threadIdx.x == 0) {
            sc->rz[(it + 1) & 1] = rzn;
            sc->pcg_iters += 1;
            if (!(pq > 0.0) || !(rzn > 0.0) || rzn <= tol2 * sc->rz_ref) sc->done_it = it + 1;
        }
    }
}
__global__ void k_axpy(float* __restrict__ t, const float* __restrict__ dl, int n3) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) t[i] += dl[i];
}

// =====================================================================================================
// persistent cooperative kernel (single rank): the whole of solveAll in one launch

struct SolveCtl {
    int num_iter, nonlinear_iter, linear_iter, early_out;
    double tol2;
    long long* prof;  // DFU_SOLVER_PROFILE: per-phase SM cycles of CTA 0 (version 3), else NULL
};

__global__ void __launch_bounds__(PTPB, 1) k_solve_persistent(Problem pb, SolveCtl ctl, Scalars* sc, unsigned* bar) {
    __shared__ double sh4[4 * (PTPB / 32)];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, gw = tid >> 5, nw = nthreads >> 5;
    const int nb = gridDim.x, n3 = 3 * pb.N;
    unsigned bar_target = 0;
    double* part4 = pb.part;  // 4 doubles per CTA
#define GRID_SYNC() grid_barrier(bar, (unsigned) nb, bar_target)

    for (int i = tid; i < n3; i += nthreads) pb.t[i] = 0.f;  // unknowns := 0 (opt_solver.cpp:192-193)
    GRID_SYNC();

    double rz_ref = -1.0, E = 0.0, E0 = 0.0;
    int pcg_total = 0, gn_total = 0;
    bool first = true, stop_all = false;

    for (int outer = 0; outer < ctl.num_iter && !stop_all; ++outer) {
        for (int gn = 0; gn < ctl.nonlinear_iter; ++gn) {
            // ---- residuals + tukey (re-weighted once per outer iteration, opt_solver.cpp:135-140) -----
            const double e2_local = phase_point_residual(pb, gn == 0, tid, nthreads);
            GRID_SYNC();
            // ---- per-node blocks: b = -J^T r, D = diag(J^T J) (+ regularisation), PCG initialisation --------
            {
                double rz = 0.0, er = 0.0;
                for (int n = gw; n < pb.N; n += nw) {
                    float ax, ay, az, ad, gx = 0.f, gy = 0.f, gz = 0.f, cnt = 0.f, e2 = 0.f;
                    node_gather_data(pb, n, lane, true, ax, ay, az, ad);
                    if (pb.wreg2 > 0.f) {
                        node_gather_reg(pb, n, lane, pb.t, gx, gy, gz, cnt, e2);
                        ax -= pb.wreg2 * gx; ay -= pb.wreg2 * gy; az -= pb.wreg2 * gz;
                        ad += pb.wreg2 * cnt;
                        e2 = warp_sum(e2);
                    }
                    ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az); ad = warp_sum(ad);
                    if (lane == 0) {
                        const float inv = ad > 0.f ? 1.f / ad : 0.f;
                        const double invd = ad > 0.f ? 1.0 / (double) ad : 0.0;
                        const float bb[3] = {ax, ay, az};
                        pb.nbuf[3 * (size_t) pb.N + n] = ad;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const size_t i = 3 * (size_t) n + c;
                            pb.nbuf[i] = bb[c]; pb.r[i] = bb[c]; pb.p[i] = bb[c] * inv; pb.dl[i] = 0.f;
                            rz += (double) bb[c] * (double) bb[c] * invd;
                        }
                        er += (double) pb.wreg2 * e2;
                    }
                }
                const D4 s = block_sum4(D4{e2_local, rz, er, 0.0}, sh4);
                if (threadIdx.x == 0) {
                    part4[4 * blockIdx.x] = s.a; part4[4 * blockIdx.x + 1] = s.b; part4[4 * blockIdx.x + 2] = s.c;
                    part4[4 * blockIdx.x + 3] = 0.0;
                }
            }
            GRID_SYNC();
            const D4 tot = sum_partials4(part4, nb, sh4);
            double rz = tot.b;
            E = tot.a + tot.c;
            if (first) {
                E0 = E;
                first = false;
            }
            if (rz_ref < 0.0) rz_ref = rz;
            const bool conv0 = !(rz > 0.0) || rz <= ctl.tol2 * rz_ref;
            if (ctl.early_out && conv0) {  // converged at this linearisation point
                if (gn == 0 && outer > 0) stop_all = true;
                GRID_SYNC();  // every CTA has read the partials before anyone overwrites them
                break;
            }
            // ---- PCG (the barrier after the point phase also separates the partials' readers and writers) ----
            if (!conv0) {
                for (int it = 0; it < ctl.linear_iter; ++it) {
                    phase_point_apply(pb, tid, nthreads);  // s4 = Theta W p
                    GRID_SYNC();
                    // p.q, r.M^-1 r, r.M^-1 q, q.M^-1 q.  r.M^-1 r is re-measured from the stored float r every
                    // iteration, so the recurrence below never drifts away from the actual residual.
                    double pq = 0.0, rr = 0.0, rmq = 0.0, qmq = 0.0;
                    for (int n = gw; n < pb.N; n += nw) {
                        float ax, ay, az, ad, gx = 0.f, gy = 0.f, gz = 0.f, cnt = 0.f, e2 = 0.f;
                        node_gather_data(pb, n, lane, false, ax, ay, az, ad);
                        if (pb.wreg2 > 0.f) {
                            node_gather_reg(pb, n, lane, pb.p, gx, gy, gz, cnt, e2);
                            ax += pb.wreg2 * gx; ay += pb.wreg2 * gy; az += pb.wreg2 * gz;
                        }
                        ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
                        if (lane == 0) {
                            const float D = pb.nbuf[3 * (size_t) pb.N + n];
                            const double inv = D > 0.f ? 1.0 / (double) D : 0.0;
                            const float qq[3] = {ax, ay, az};
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                const size_t i = 3 * (size_t) n + c;
                                const double ri = (double) pb.r[i];
                                pb.q[i] = qq[c];
                                pq += (double) pb.p[i] * qq[c];
                                rr += ri * ri * inv;
                                rmq += ri * qq[c] * inv;
                                qmq += (double) qq[c] * qq[c] * inv;
                            }
                        }
                    }
                    {
                        const D4 s = block_sum4(D4{pq, rr, rmq, qmq}, sh4);
                        if (threadIdx.x == 0) {
                            part4[4 * blockIdx.x] = s.a; part4[4 * blockIdx.x + 1] = s.b; part4[4 * blockIdx.x + 2] = s.c;
                            part4[4 * blockIdx.x + 3] = s.d;
                        }
                    }
                    GRID_SYNC();
                    const D4 g = sum_partials4(part4, nb, sh4);
                    pq = g.a; rz = g.b; rmq = g.c; qmq = g.d;
                    ++pcg_total;
                    if (!(pq > 0.0) || !(rz > 0.0)) break;
                    // r' = r - alpha q, z' = M^-1 r'  =>  r'.z' = r.z - 2 alpha r.M^-1 q + alpha^2 q.M^-1 q
                    const double alpha = rz / pq;
                    double rzn = rz - 2.0 * alpha * rmq + alpha * alpha * qmq;
                    if (!(rzn > 0.0)) rzn = 0.0;
                    const float af = (float) alpha, bf = (float) (rzn / rz);
                    for (int i = tid; i < n3; i += nthreads) {
                        const float D = pb.nbuf[3 * (size_t) pb.N + i / 3];
                        const float inv = D > 0.f ? 1.f / D : 0.f;
                        const float p = pb.p[i];
                        pb.dl[i] = __fmaf_rn(af, p, pb.dl[i]);
                        const float r = __fmaf_rn(-af, pb.q[i], pb.r[i]);
                        pb.r[i] = r;
                        pb.p[i] = __fmaf_rn(bf, p, r * inv);
                    }
                    rz = rzn;
                    GRID_SYNC();
                    if (!(rz > 0.0) || rz <= ctl.tol2 * rz_ref) break;
                }
            }
            GRID_SYNC();  // (also covers the PCG exits that left without a barrier after reading the partials)
            for (int i = tid; i < n3; i += nthreads) pb.t[i] += pb.dl[i];
            ++gn_total;
            GRID_SYNC();
        }
    }
    // ---- final energy at the solution, Tukey weights of the last outer iteration -------------------------
    {
        const double e2 = phase_point_residual(pb, first, tid, nthreads);  // no GN step ran: weights at t = 0
        double er = 0.0;
        if (pb.wreg2 > 0.f)
            for (int n = gw; n < pb.N; n += nw) {
                float gx, gy, gz, cnt, r2;
                node_gather_reg(pb, n, lane, pb.t, gx, gy, gz, cnt, r2);
                r2 = warp_sum(r2);
                if (lane == 0) er += (double) pb.wreg2 * r2;
            }
        const D4 s = block_sum4(D4{e2, er, 0.0, 0.0}, sh4);
        if (threadIdx.x == 0) {
            part4[4 * blockIdx.x] = s.a; part4[4 * blockIdx.x + 1] = s.b; part4[4 * blockIdx.x + 2] = 0.0;
            part4[4 * blockIdx.x + 3] = 0.0;
        }
    }
    GRID_SYNC();
    {
        const D4 g = sum_partials4(part4, nb, sh4);
        E = g.a + g.b;
    }
    if (tid == 0) {
        sc->E = E;
        sc->E0 = first ? E : E0;
        sc->rz_ref = rz_ref;
        sc->pcg_iters = pcg_total;
        sc->gn_steps = gn_total;
        sc->first = 0;
    }
#undef GRID_SYNC
}

// =====================================================================================================
// persistent kernel, version 2: same loop, iteration-invariant data in REGISTERS.
//
// The graph is fixed during a solve and the assignment of points to threads / nodes to warps is static, so:
//   * every thread keeps the 8 neighbour ids + 8 weights + (live - canon) + tukey weight of its point;
//   * every warp keeps, for each of its (at most P2_NPW) nodes, 8 transposed-list entries per lane (256 per node;
//     longer lists finish from L2) and one regularisation edge per lane; lane 0 keeps the node's D, r, p, q, delta, t.
// Only p, t (gathered by other CTAs), s4 and the per-CTA partial sums go through memory, so a PCG iteration is three
// grid barriers with ONE L2 round trip each: gather p | gather s4 (+ p for the regulariser) | read the partials.
constexpr int P2_NPW = 2;   // nodes per warp held in registers (N <= 2 * 148 * 16 = 4736; more nodes: version 1)
constexpr int P2_LE = 8;    // transposed-list entries per lane per node held in registers

__global__ void __launch_bounds__(PTPB, 1) k_solve_persistent2(Problem pb, SolveCtl ctl, Scalars* sc, unsigned* bar) {
    __shared__ double sh4[4 * (PTPB / 32)];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, gw = tid >> 5, nw = nthreads >> 5;
    const int nb = gridDim.x;
    unsigned bar_target = 0;
    double* part4 = pb.part;
#define GRID_SYNC() grid_barrier(bar, (unsigned) nb, bar_target)

    // ---- my point -------------------------------------------------------------------------------------------
    const bool has_pt = tid < pb.P;
    int my_nb[8];
    float my_w[8], my_th = 0.f, my_d0 = 0.f, my_d1 = 0.f, my_d2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        my_nb[k] = 0;
        my_w[k] = 0.f;
    }
    if (has_pt) {
        load8(pb.nbr, pb.wts, tid, my_nb, my_w);
        my_d0 = pb.dvec[3 * (size_t) tid]; my_d1 = pb.dvec[3 * (size_t) tid + 1]; my_d2 = pb.dvec[3 * (size_t) tid + 2];
    }
    // ---- my warp's nodes --------------------------------------------------------------------------------------
    int node[P2_NPW], l_hi[P2_NPW], l_lo[P2_NPW];  // node id (-1: none), list range
    int ev[P2_NPW][P2_LE];                         // list entries of this lane: point ids (-1: none)
    float ew[P2_NPW][P2_LE];                       //                             weights
    int redge[P2_NPW];                             // this lane's regularisation edge target (-1: none)
    bool rout[P2_NPW];                             // ... is an out-edge (counts for the energy)
    int rin_lo[P2_NPW], rin_n[P2_NPW];
    float nD[P2_NPW], nr[P2_NPW][3], np_[P2_NPW][3], nq[P2_NPW][3], ndl[P2_NPW][3], nt[P2_NPW][3];  // used by lane 0
#pragma unroll
    for (int sidx = 0; sidx < P2_NPW; ++sidx) {
        const int n = gw + sidx * nw;
        node[sidx] = n < pb.N ? n : -1;
        l_lo[sidx] = l_hi[sidx] = 0;
        redge[sidx] = -1;
        rout[sidx] = false;
        rin_lo[sidx] = rin_n[sidx] = 0;
        nD[sidx] = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) nr[sidx][c] = np_[sidx][c] = nq[sidx][c] = ndl[sidx][c] = nt[sidx][c] = 0.f;
#pragma unroll
        for (int u = 0; u < P2_LE; ++u) {
            ev[sidx][u] = -1;
            ew[sidx][u] = 0.f;
        }
        if (n < pb.N) {
            l_lo[sidx] = pb.tptr[n];
            l_hi[sidx] = pb.tptr[n + 1];
#pragma unroll
            for (int u = 0; u < P2_LE; ++u) {
                const int j = l_lo[sidx] + lane + 32 * u;
                if (j < l_hi[sidx]) {
                    ev[sidx][u] = pb.tv[j];
                    ew[sidx][u] = pb.tw[j];
                }
            }
            rin_lo[sidx] = pb.rin_ptr[n];
            rin_n[sidx] = pb.rin_ptr[n + 1] - rin_lo[sidx];
            if (lane < 8) {
                redge[sidx] = pb.nnbr[(size_t) n * 8 + lane];
                rout[sidx] = true;
            } else if (lane - 8 < rin_n[sidx]) {
                redge[sidx] = pb.rin[rin_lo[sidx] + lane - 8];
            }
            if (redge[sidx] == n) redge[sidx] = -1;
        }
    }
    for (int n = gw + P2_NPW * nw; n < pb.N; n += nw)  // (never taken when N <= P2_NPW * warps; kept for safety)
        if (lane == 0) pb.t[3 * (size_t) n] = pb.t[3 * (size_t) n + 1] = pb.t[3 * (size_t) n + 2] = 0.f;
#pragma unroll
    for (int sidx = 0; sidx < P2_NPW; ++sidx)
        if (node[sidx] >= 0 && lane == 0)  // unknowns := 0 (opt_solver.cpp:192-193)
            pb.t[3 * (size_t) node[sidx]] = pb.t[3 * (size_t) node[sidx] + 1] = pb.t[3 * (size_t) node[sidx] + 2] = 0.f;
    GRID_SYNC();

    // sum_k w_k x[n_k] for my point
    auto my_gather = [&](const float* x, float& sx, float& sy, float& sz) {
        sx = sy = sz = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float* xk = x + 3 * (size_t) my_nb[k];
            sx = __fmaf_rn(my_w[k], xk[0], sx);
            sy = __fmaf_rn(my_w[k], xk[1], sy);
            sz = __fmaf_rn(my_w[k], xk[2], sz);
        }
    };
    auto residual_phase = [&](bool update_tukey) -> double {
        double e2 = 0.0;
        if (has_pt) {
            float sx, sy, sz;
            my_gather(pb.t, sx, sy, sz);
            const float ex = my_d0 - sx, ey = my_d1 - sy, ez = my_d2 - sz;
            if (update_tukey) my_th = tukey_biweight(pb.tukey_offset, pb.psi_data, ex, ey, ez);
            pb.s4[tid] = make_float4(my_th * ex, my_th * ey, my_th * ez, my_th);
            e2 = (double) my_th * ((double) ex * ex + (double) ey * ey + (double) ez * ez);
        }
        for (int v = tid + nthreads; v < pb.P; v += nthreads) {  // more points than threads: the rest from L2
            float sx, sy, sz;
            point_gather(pb, v, pb.t, sx, sy, sz);
            const float ex = pb.dvec[3 * (size_t) v] - sx, ey = pb.dvec[3 * (size_t) v + 1] - sy, ez = pb.dvec[3 * (size_t) v + 2] - sz;
            float th;
            if (update_tukey) {
                th = tukey_biweight(pb.tukey_offset, pb.psi_data, ex, ey, ez);
                pb.theta[v] = th;
            } else {
                th = pb.theta[v];
            }
            pb.s4[v] = make_float4(th * ex, th * ey, th * ez, th);
            e2 += (double) th * ((double) ex * ex + (double) ey * ey + (double) ez * ez);
        }
        return e2;
    };
    // data gather of slot sidx: registers first, entries beyond 32*P2_LE from L2 (same order as version 1)
    auto gather_data = [&](int sidx, bool with_diag, float& ax, float& ay, float& az, float& ad) {
        ax = ay = az = ad = 0.f;
        float4 sv[P2_LE];
#pragma unroll
        for (int u = 0; u < P2_LE; ++u) sv[u] = ev[sidx][u] >= 0 ? pb.s4[ev[sidx][u]] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < P2_LE; ++u) {
            if (ev[sidx][u] < 0) continue;
            const float w = ew[sidx][u];
            ax = __fmaf_rn(w, sv[u].x, ax);
            ay = __fmaf_rn(w, sv[u].y, ay);
            az = __fmaf_rn(w, sv[u].z, az);
            if (with_diag) ad = __fmaf_rn(w * w, sv[u].w, ad);
        }
        for (int j = l_lo[sidx] + 32 * P2_LE + lane; j < l_hi[sidx]; j += 32) {
            const float w = pb.tw[j];
            const float4 s4v = pb.s4[pb.tv[j]];
            ax = __fmaf_rn(w, s4v.x, ax);
            ay = __fmaf_rn(w, s4v.y, ay);
            az = __fmaf_rn(w, s4v.z, az);
            if (with_diag) ad = __fmaf_rn(w * w, s4v.w, ad);
        }
    };
    // regularisation gather of slot sidx on published vector x; own value broadcast from lane 0's registers
    auto gather_reg = [&](int sidx, const float* x, float o0, float o1, float o2, float& gx, float& gy, float& gz, float& cnt, float& e2) {
        gx = gy = gz = cnt = e2 = 0.f;
        const float x0 = __shfl_sync(0xffffffffu, o0, 0), x1 = __shfl_sync(0xffffffffu, o1, 0), x2 = __shfl_sync(0xffffffffu, o2, 0);
        if (redge[sidx] >= 0) {
            const int m = redge[sidx];
            const float d0 = x0 - x[3 * (size_t) m], d1 = x1 - x[3 * (size_t) m + 1], d2 = x2 - x[3 * (size_t) m + 2];
            gx = d0; gy = d1; gz = d2;
            cnt = 1.f;
            if (rout[sidx]) e2 = d0 * d0 + d1 * d1 + d2 * d2;
        }
        for (int j = 24 + lane; j < rin_n[sidx]; j += 32) {  // in-edges beyond the 24 held in registers
            const int m = pb.rin[rin_lo[sidx] + j];
            if (m == node[sidx]) continue;
            gx += x0 - x[3 * (size_t) m]; gy += x1 - x[3 * (size_t) m + 1]; gz += x2 - x[3 * (size_t) m + 2];
            cnt += 1.f;
        }
    };

    double rz_ref = -1.0, E = 0.0, E0 = 0.0;
    int pcg_total = 0, gn_total = 0;
    bool first = true, stop_all = false;

    for (int outer = 0; outer < ctl.num_iter && !stop_all; ++outer) {
        for (int gn = 0; gn < ctl.nonlinear_iter; ++gn) {
            const double e2_local = residual_phase(gn == 0);
            GRID_SYNC();
            {   // per-node blocks b = -J^T r, D = diag(J^T J) (+ regularisation), PCG initialisation
                double rz = 0.0, er = 0.0;
#pragma unroll
                for (int sidx = 0; sidx < P2_NPW; ++sidx) {
                    if (node[sidx] < 0) continue;  // warp-uniform
                    float ax, ay, az, ad, gx = 0.f, gy = 0.f, gz = 0.f, cnt = 0.f, e2 = 0.f;
                    gather_data(sidx, true, ax, ay, az, ad);
                    if (pb.wreg2 > 0.f) {
                        gather_reg(sidx, pb.t, nt[sidx][0], nt[sidx][1], nt[sidx][2], gx, gy, gz, cnt, e2);
                        ax -= pb.wreg2 * gx; ay -= pb.wreg2 * gy; az -= pb.wreg2 * gz;
                        ad += pb.wreg2 * cnt;
                        e2 = warp_sum(e2);
                    }
                    ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az); ad = warp_sum(ad);
                    if (lane == 0) {
                        const float inv = ad > 0.f ? 1.f / ad : 0.f;
                        const double invd = ad > 0.f ? 1.0 / (double) ad : 0.0;
                        const float bb[3] = {ax, ay, az};
                        nD[sidx] = ad;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            nr[sidx][c] = bb[c];
                            np_[sidx][c] = bb[c] * inv;
                            ndl[sidx][c] = 0.f;
                            pb.p[3 * (size_t) node[sidx] + c] = np_[sidx][c];
                            rz += (double) bb[c] * (double) bb[c] * invd;
                        }
                        er += (double) pb.wreg2 * e2;
                    }
                }
                for (int n = gw + P2_NPW * nw; n < pb.N; n += nw) {  // nodes beyond the register slots: not supported here
                }
                const D4 s = block_sum4(D4{e2_local, rz, er, 0.0}, sh4);
                if (threadIdx.x == 0) {
                    part4[4 * blockIdx.x] = s.a; part4[4 * blockIdx.x + 1] = s.b; part4[4 * blockIdx.x + 2] = s.c;
                    part4[4 * blockIdx.x + 3] = 0.0;
                }
            }
            GRID_SYNC();
            const D4 tot = sum_partials4(part4, nb, sh4);
            double rz = tot.b;
            E = tot.a + tot.c;
            if (first) {
                E0 = E;
                first = false;
            }
            if (rz_ref < 0.0) rz_ref = rz;
            const bool conv0 = !(rz > 0.0) || rz <= ctl.tol2 * rz_ref;
            if (ctl.early_out && conv0) {  // converged at this linearisation point
                if (gn == 0 && outer > 0) stop_all = true;
                GRID_SYNC();  // every CTA has read the partials before anyone overwrites them
                break;
            }
            if (!conv0) {
                for (int it = 0; it < ctl.linear_iter; ++it) {
                    // ---- s4 = Theta W p ----
                    if (has_pt) {
                        float sx = 0.f, sy = 0.f, sz = 0.f;
                        if (my_th != 0.f) my_gather(pb.p, sx, sy, sz);
                        pb.s4[tid] = make_float4(my_th * sx, my_th * sy, my_th * sz, my_th);
                    }
                    for (int v = tid + nthreads; v < pb.P; v += nthreads) {
                        const float th = pb.theta[v];
                        float sx = 0.f, sy = 0.f, sz = 0.f;
                        if (th != 0.f) point_gather(pb, v, pb.p, sx, sy, sz);
                        pb.s4[v] = make_float4(th * sx, th * sy, th * sz, th);
                    }
                    GRID_SYNC();
                    // ---- q = W^T s4 + w_reg^2 L p for my nodes; p.q, r.M^-1 r, r.M^-1 q, q.M^-1 q ----
                    double pq = 0.0, rr = 0.0, rmq = 0.0, qmq = 0.0;
#pragma unroll
                    for (int sidx = 0; sidx < P2_NPW; ++sidx) {
                        if (node[sidx] < 0) continue;
                        float ax, ay, az, ad, gx = 0.f, gy = 0.f, gz = 0.f, cnt = 0.f, e2 = 0.f;
                        gather_data(sidx, false, ax, ay, az, ad);
                        if (pb.wreg2 > 0.f) {
                            gather_reg(sidx, pb.p, np_[sidx][0], np_[sidx][1], np_[sidx][2], gx, gy, gz, cnt, e2);
                            ax += pb.wreg2 * gx; ay += pb.wreg2 * gy; az += pb.wreg2 * gz;
                        }
                        ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
                        if (lane == 0) {
                            const double inv = nD[sidx] > 0.f ? 1.0 / (double) nD[sidx] : 0.0;
                            const float qq[3] = {ax, ay, az};
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                const double ri = (double) nr[sidx][c];
                                nq[sidx][c] = qq[c];
                                pq += (double) np_[sidx][c] * qq[c];
                                rr += ri * ri * inv;
                                rmq += ri * qq[c] * inv;
                                qmq += (double) qq[c] * qq[c] * inv;
                            }
                        }
                    }
                    {
                        const D4 s = block_sum4(D4{pq, rr, rmq, qmq}, sh4);
                        if (threadIdx.x == 0) {
                            part4[4 * blockIdx.x] = s.a; part4[4 * blockIdx.x + 1] = s.b; part4[4 * blockIdx.x + 2] = s.c;
                            part4[4 * blockIdx.x + 3] = s.d;
                        }
                    }
                    GRID_SYNC();
                    const D4 g = sum_partials4(part4, nb, sh4);
                    pq = g.a; rz = g.b; rmq = g.c; qmq = g.d;
                    ++pcg_total;
                    if (!(pq > 0.0) || !(rz > 0.0)) break;
                    // r' = r - alpha q, z' = M^-1 r'  =>  r'.z' = r.z - 2 alpha r.M^-1 q + alpha^2 q.M^-1 q
                    const double alpha = rz / pq;
                    double rzn = rz - 2.0 * alpha * rmq + alpha * alpha * qmq;
                    if (!(rzn > 0.0)) rzn = 0.0;
                    const float af = (float) alpha, bf = (float) (rzn / rz);
                    if (lane == 0) {
#pragma unroll
                        for (int sidx = 0; sidx < P2_NPW; ++sidx) {
                            if (node[sidx] < 0) continue;
                            const float inv = nD[sidx] > 0.f ? 1.f / nD[sidx] : 0.f;
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                const float p = np_[sidx][c];
                                ndl[sidx][c] = __fmaf_rn(af, p, ndl[sidx][c]);
                                const float r = __fmaf_rn(-af, nq[sidx][c], nr[sidx][c]);
                                nr[sidx][c] = r;
                                np_[sidx][c] = __fmaf_rn(bf, p, r * inv);
                                pb.p[3 * (size_t) node[sidx] + c] = np_[sidx][c];
                            }
                        }
                    }
                    rz = rzn;
                    GRID_SYNC();
                    if (!(rz > 0.0) || rz <= ctl.tol2 * rz_ref) break;
                }
            }
            GRID_SYNC();  // (also covers the PCG exits that left without a barrier after reading the partials)
            if (lane == 0) {
#pragma unroll
                for (int sidx = 0; sidx < P2_NPW; ++sidx) {
                    if (node[sidx] < 0) continue;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        nt[sidx][c] += ndl[sidx][c];
                        ndl[sidx][c] = 0.f;
                        pb.t[3 * (size_t) node[sidx] + c] = nt[sidx][c];
                    }
                }
            }
            ++gn_total;
            GRID_SYNC();
        }
    }
    // ---- final energy at the solution, Tukey weights of the last outer iteration -------------------------
    {
        const double e2 = residual_phase(first);  // no GN step ran: weights at t = 0
        double er = 0.0;
        if (pb.wreg2 > 0.f) {
#pragma unroll
            for (int sidx = 0; sidx < P2_NPW; ++sidx) {
                if (node[sidx] < 0) continue;
                float gx, gy, gz, cnt, r2;
                gather_reg(sidx, pb.t, nt[sidx][0], nt[sidx][1], nt[sidx][2], gx, gy, gz, cnt, r2);
                r2 = warp_sum(r2);
                if (lane == 0) er += (double) pb.wreg2 * r2;
            }
        }
        const D4 s = block_sum4(D4{e2, er, 0.0, 0.0}, sh4);
        if (threadIdx.x == 0) {
            part4[4 * blockIdx.x] = s.a; part4[4 * blockIdx.x + 1] = s.b; part4[4 * blockIdx.x + 2] = 0.0;
            part4[4 * blockIdx.x + 3] = 0.0;
        }
    }
    if (has_pt) pb.theta[tid] = my_th;  // keep the global copy coherent
    GRID_SYNC();
    {
        const D4 g = sum_partials4(part4, nb, sh4);
        E = g.a + g.b;
    }
    if (tid == 0) {
        sc->E = E;
        sc->E0 = first ? E : E0;
        sc->rz_ref = rz_ref;
        sc->pcg_iters = pcg_total;
        sc->gn_steps = gn_total;
        sc->first = 0;
    }
#undef GRID_SYNC
}

// =====================================================================================================
// persistent kernel, version 3: explicit normal matrix + pipelined preconditioned CG.
//
// Versions 1/2 apply A = W^T Theta W + w_reg^2 L matrix-free: every PCG iteration walks the 8P graph edges twice
// (points, then nodes) with 3 grid barriers.  But A is tiny -- N rows with a few dozen non-zeros (nodes that share a
// surface point or a regularisation edge) -- and fixed between re-weightings.  So it is assembled once per GN step
// (one warp per row, gathers over the transposed graph, 64-bit fixed-point accumulation in shared memory: integer
// adds are associative, hence bit-reproducible whatever the arrival order) and a PCG iteration becomes one sparse
// row product per node.  The iteration itself is the pipelined preconditioned CG of Ghysels & Vanroose
// (Parallel Computing 40, 2014, alg. 4; same iterates as textbook PCG in exact arithmetic): the two dot products
// (r,u), (w,u) and the matrix product n = A M^-1 w of one iteration do not depend on each other, so they share ONE
// grid barrier.  u = M^-1 r and m = M^-1 w are recomputed from r and w (M is diagonal), which removes two of the
// recurrences of the published algorithm.  All per-row vectors are touched only by the warp that owns the row.
DFU_DEV void spmv_row(const Pattern& pt, int off, int len, int lane, const float4* __restrict__ x, float& ax, float& ay, float& az) {
    ax = ay = az = 0.f;
    for (int j = lane; j < len; j += 32) {
        const float v = pt.vals[off + j];
        const float4 m = x[pt.col[off + j]];
        ax = __fmaf_rn(v, m.x, ax);
        ay = __fmaf_rn(v, m.y, ay);
        az = __fmaf_rn(v, m.z, az);
    }
    ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
}

__global__ void __launch_bounds__(PTPB, 1) k_solve_persistent3(Problem pb, Pattern pt, SolveCtl ctl, Scalars* sc, unsigned* bar) {
    __shared__ double sh4[4 * (PTPB / 32)];
    __shared__ unsigned long long acc_sm[(PTPB / 32) * ACC_W];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, gw = tid >> 5, nw = nthreads >> 5;
    const int nb = gridDim.x, N = pb.N;
    // every warp owns a contiguous block of rows (at most 32: N <= 65535 and >= 2368 resident warps): warp-per-row for the
    // row products, lane-per-row (coalesced float4 accesses) for the vector updates
    const int R = (N + nw - 1) / nw;
    const int row0 = min(N, gw * R), row1 = min(N, row0 + R);
    unsigned bar_target = 0;
    unsigned long long* acc = acc_sm + (threadIdx.x >> 5) * ACC_W;
    float4 *S_r = pt.st, *S_w = pt.st + N, *S_z = pt.st + 2 * (size_t) N, *S_s = pt.st + 3 * (size_t) N,
           *S_p = pt.st + 4 * (size_t) N, *S_x = pt.st + 5 * (size_t) N;
#define GRID_SYNC() grid_barrier(bar, (unsigned) nb, bar_target)
#define PART(buf) (pb.part + (size_t) (buf) * 2 * MAX_PARTIALS)
    long long t_prev = clock64();
#define PROF(k)                                               \
    do {                                                      \
        if (ctl.prof && tid == 0) {                           \
            const long long t_now = clock64();                \
            ctl.prof[k] += t_now - t_prev;                    \
            t_prev = t_now;                                   \
        }                                                     \
    } while (0)

    for (int i = tid; i < 3 * N; i += nthreads) pb.t[i] = 0.f;  // unknowns := 0 (opt_solver.cpp:192-193)
    GRID_SYNC();

    // y = A x for all rows of this warp, four rows in flight (their index / value / gather loads are issued together);
    // the lane that owns row n (lane == n - row0) receives the result
    auto spmv_block = [&](const float4* __restrict__ x, float& rx, float& ry, float& rz_) {
        rx = ry = rz_ = 0.f;
        for (int n0 = row0; n0 < row1; n0 += 4) {
            int off[4], len[4], c[4][2];
            float v[4][2];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool ok = n0 + q < row1;
                off[q] = ok ? pt.rowptr[n0 + q] : 0;
                len[q] = ok ? pt.rowlen[n0 + q] : 0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int j = lane + 32 * u;
                    c[q][u] = j < len[q] ? pt.col[off[q] + j] : -1;
                    v[q][u] = j < len[q] ? pt.vals[off[q] + j] : 0.f;
                }
            float ax[4], ay[4], az[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                ax[q] = ay[q] = az[q] = 0.f;
#pragma unroll
                for (int u = 0; u < 2; ++u)
                    if (c[q][u] >= 0) {
                        const float4 m = __ldcg(x + c[q][u]);
                        ax[q] = __fmaf_rn(v[q][u], m.x, ax[q]);
                        ay[q] = __fmaf_rn(v[q][u], m.y, ay[q]);
                        az[q] = __fmaf_rn(v[q][u], m.z, az[q]);
                    }
                for (int j = 64 + lane; j < len[q]; j += 32) {  // rows longer than 64 entries: the rest
                    const float vv = pt.vals[off[q] + j];
                    const float4 m = __ldcg(x + pt.col[off[q] + j]);
                    ax[q] = __fmaf_rn(vv, m.x, ax[q]);
                    ay[q] = __fmaf_rn(vv, m.y, ay[q]);
                    az[q] = __fmaf_rn(vv, m.z, az[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                ax[q] = warp_sum(ax[q]); ay[q] = warp_sum(ay[q]); az[q] = warp_sum(az[q]);
                if (lane == n0 + q - row0) {
                    rx = ax[q]; ry = ay[q]; rz_ = az[q];
                }
            }
        }
    };

    double rz_ref = -1.0, E = 0.0, E0 = 0.0;
    int pcg_total = 0, gn_total = 0;
    bool first = true, stop_all = false;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int outer = 0; outer < ctl.num_iter && !stop_all; ++outer) {
        for (int gn = 0; gn < ctl.nonlinear_iter; ++gn) {
            // ---- residuals + tukey (re-weighted once per outer iteration, opt_solver.cpp:135-140) -----
            PROF(0);
            const double e2_local = phase_point_residual(pb, gn == 0, tid, nthreads);
            PROF(1);
            GRID_SYNC();
            PROF(2);
            // ---- per row: b = -J^T r (+ regularisation), the row of A, D = A_nn, PCG start r = b, u = M^-1 b, x = 0 ----
            {
                double rz = 0.0, er = 0.0;
                // (strided, not the contiguous blocks: neighbouring rows are equally heavy, and this phase walks whole
                //  (point, weight) lists -- any warp may prepare any row, its outputs all go to memory)
                for (int n = gw; n < N; n += nw) {
                    float ax, ay, az, gx = 0.f, gy = 0.f, gz = 0.f, cnt = 0.f, e2 = 0.f;
                    node_gather_data_fixed(pb, n, lane, ax, ay, az);  // already summed over the warp
                    if (pb.wreg2 > 0.f) {
                        node_gather_reg(pb, n, lane, pb.t, gx, gy, gz, cnt, e2);
                        gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
                        ax -= pb.wreg2 * gx; ay -= pb.wreg2 * gy; az -= pb.wreg2 * gz;
                        e2 = warp_sum(e2);
                    }
                    const int off = pt.rowptr[n], len = pt.rowlen[n];
                    PROF(3);
                    if (gn == 0) {  // theta changed: data part of the row, ACC_W columns per pass
                        const int lo = pb.tptr[n], hi = pb.tptr[n + 1];
                        for (int c0 = 0; c0 < len; c0 += ACC_W) {
                            for (int j = lane; j < ACC_W; j += 32) acc[j] = 0ull;
                            __syncwarp();
                            for (int e = lo + lane; e < hi; e += 32) {
                                const int v = pb.tv[e];
                                const float th = pb.theta[v];
                                if (th == 0.f) continue;
                                const float c = th * pb.tw[e];
                                const float4 w0 = *(reinterpret_cast<const float4*>(pb.wts) + 2 * (size_t) v);
                                const float4 w1 = *(reinterpret_cast<const float4*>(pb.wts) + 2 * (size_t) v + 1);
                                const uint4 sl = pt.tslot[e];
                                const float wk[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                                const unsigned sk[8] = {sl.x & 0xffffu, sl.x >> 16, sl.y & 0xffffu, sl.y >> 16,
                                                        sl.z & 0xffffu, sl.z >> 16, sl.w & 0xffffu, sl.w >> 16};
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    const unsigned sidx = sk[k] - (unsigned) c0;
                                    if (sidx < (unsigned) ACC_W)
                                        atomicAdd(&acc[sidx], (unsigned long long) __float2ll_rn(c * wk[k] * FIX_SCALE));
                                }
                            }
                            __syncwarp();
                            for (int j = lane; j < ACC_W && c0 + j < len; j += 32)
                                pt.vals[off + c0 + j] = pt.areg[off + c0 + j] + (float) ((double) (long long) acc[j] * FIX_INV);
                            __syncwarp();
                        }
                    }
                    __syncwarp();
                    PROF(4);
                    const float D = pt.vals[off + pt.dslot[n]];
                    if (lane == 0) {
                        const float inv = D > 0.f ? 1.f / D : 0.f;
                        const double invd = D > 0.f ? 1.0 / (double) D : 0.0;
                        pb.nbuf[3 * (size_t) N + n] = D;
                        pb.nbuf[3 * (size_t) n] = ax; pb.nbuf[3 * (size_t) n + 1] = ay; pb.nbuf[3 * (size_t) n + 2] = az;
                        S_r[n] = make_float4(ax, ay, az, 0.f);
                        S_x[n] = zero4;
                        pt.exch[n] = make_float4(ax * inv, ay * inv, az * inv, 0.f);
                        rz += ((double) ax * ax + (double) ay * ay + (double) az * az) * invd;
                        er += (double) pb.wreg2 * e2;
                    }
                }
                const D4 s = block_sum4(D4{e2_local, rz, er, 0.0}, sh4);
                if (threadIdx.x == 0) {
                    double* part4 = PART(0);
                    part4[4 * blockIdx.x] = s.a; part4[4 * blockIdx.x + 1] = s.b; part4[4 * blockIdx.x + 2] = s.c;
                    part4[4 * blockIdx.x + 3] = 0.0;
                }
            }
            PROF(5);
            GRID_SYNC();
            PROF(6);
            const D4 tot = sum_partials4(PART(0), nb, sh4);
            const double rz0 = tot.b;
            E = tot.a + tot.c;
            if (first) {
                E0 = E;
                first = false;
            }
            if (rz_ref < 0.0) rz_ref = rz0;
            const bool conv0 = !(rz0 > 0.0) || rz0 <= ctl.tol2 * rz_ref;
            if (ctl.early_out && conv0) {  // converged at this linearisation point
                if (gn == 0 && outer > 0) stop_all = true;
                GRID_SYNC();  // every CTA has read the partials before anyone overwrites them
                break;
            }
            PROF(7);
            if (!conv0) {
                // w0 = A u0; z = s = p = 0
                {
                    float wx, wy, wz;
                    spmv_block(pt.exch, wx, wy, wz);
                    const int n = row0 + lane;
                    if (n < row1) {
                        S_w[n] = make_float4(wx, wy, wz, 0.f);
                        S_z[n] = zero4; S_s[n] = zero4; S_p[n] = zero4;
                    }
                }
                double gamma_prev = 0.0, alpha_prev = 0.0;
                PROF(8);
                for (int it = 0; it < ctl.linear_iter; ++it) {
                    const int buf = (it + 1) & 1;
                    float4* ex = pt.exch + (size_t) buf * N;
                    // (r,u), (w,u) with u = M^-1 r; m = M^-1 w goes to the exchange buffer
                    double g = 0.0, d = 0.0;
                    {
                        const int n = row0 + lane;  // one row per lane
                        if (n < row1) {
                            const float D = pb.nbuf[3 * (size_t) N + n];
                            const float inv = D > 0.f ? 1.f / D : 0.f;
                            const double invd = D > 0.f ? 1.0 / (double) D : 0.0;
                            const float4 r = S_r[n], w = S_w[n];
                            g += ((double) r.x * r.x + (double) r.y * r.y + (double) r.z * r.z) * invd;
                            d += ((double) w.x * r.x + (double) w.y * r.y + (double) w.z * r.z) * invd;
                            ex[n] = make_float4(w.x * inv, w.y * inv, w.z * inv, 0.f);
                        }
                    }
                    {
                        const D4 s = block_sum4(D4{g, d, 0.0, 0.0}, sh4);
                        if (threadIdx.x == 0) {
                            double* part4 = PART(buf);
                            part4[4 * blockIdx.x] = s.a; part4[4 * blockIdx.x + 1] = s.b;
                            part4[4 * blockIdx.x + 2] = 0.0; part4[4 * blockIdx.x + 3] = 0.0;
                        }
                    }
                    PROF(9);
                    GRID_SYNC();
                    PROF(10);
                    const D4 gd = sum_partials4(PART(buf), nb, sh4);
                    const double gamma = gd.a, delta = gd.b;
                    if (!(gamma > 0.0) || (it > 0 && gamma <= ctl.tol2 * rz_ref)) break;
                    const double beta = it > 0 ? gamma / gamma_prev : 0.0;
                    const double denom = it > 0 ? delta - beta * gamma / alpha_prev : delta;
                    if (!(denom > 0.0)) break;
                    const double alpha = gamma / denom;
                    const float af = (float) alpha, bf = (float) beta;
                    const bool last = it + 1 >= ctl.linear_iter;
                    PROF(11);
                    {
                        float nx = 0.f, ny = 0.f, nz = 0.f;
                        if (!last) spmv_block(ex, nx, ny, nz);  // n = A m
                        const int n = row0 + lane;
                        if (n < row1) {
                            const float D = pb.nbuf[3 * (size_t) N + n];
                            const float inv = D > 0.f ? 1.f / D : 0.f;
                            float4 r = S_r[n], w = S_w[n], z = S_z[n], sv = S_s[n], p = S_p[n], x = S_x[n];
                            z.x = __fmaf_rn(bf, z.x, nx); z.y = __fmaf_rn(bf, z.y, ny); z.z = __fmaf_rn(bf, z.z, nz);
                            sv.x = __fmaf_rn(bf, sv.x, w.x); sv.y = __fmaf_rn(bf, sv.y, w.y); sv.z = __fmaf_rn(bf, sv.z, w.z);
                            p.x = __fmaf_rn(bf, p.x, r.x * inv); p.y = __fmaf_rn(bf, p.y, r.y * inv); p.z = __fmaf_rn(bf, p.z, r.z * inv);
                            x.x = __fmaf_rn(af, p.x, x.x); x.y = __fmaf_rn(af, p.y, x.y); x.z = __fmaf_rn(af, p.z, x.z);
                            r.x = __fmaf_rn(-af, sv.x, r.x); r.y = __fmaf_rn(-af, sv.y, r.y); r.z = __fmaf_rn(-af, sv.z, r.z);
                            w.x = __fmaf_rn(-af, z.x, w.x); w.y = __fmaf_rn(-af, z.y, w.y); w.z = __fmaf_rn(-af, z.z, w.z);
                            S_r[n] = r; S_w[n] = w; S_z[n] = z; S_s[n] = sv; S_p[n] = p; S_x[n] = x;
                        }
                    }
                    PROF(12);
                    ++pcg_total;
                    gamma_prev = gamma;
                    alpha_prev = alpha;
                }
            }
            // t += x (row-local), then everybody needs the new t
            {
                const int n = row0 + lane;
                if (n < row1) {
                    const float4 x = S_x[n];
                    pb.t[3 * (size_t) n] += x.x; pb.t[3 * (size_t) n + 1] += x.y; pb.t[3 * (size_t) n + 2] += x.z;
                }
            }
            ++gn_total;
            PROF(13);
            GRID_SYNC();
            PROF(14);
        }
    }
    // ---- final energy at the solution, Tukey weights of the last outer iteration -------------------------
    {
        const double e2 = phase_point_residual(pb, first, tid, nthreads);  // no GN step ran: weights at t = 0
        double er = 0.0;
        if (pb.wreg2 > 0.f)
            for (int n = row0; n < row1; ++n) {
                float gx, gy, gz, cnt, r2;
                node_gather_reg(pb, n, lane, pb.t, gx, gy, gz, cnt, r2);
                r2 = warp_sum(r2);
                if (lane == 0) er += (double) pb.wreg2 * r2;
            }
        const D4 s = block_sum4(D4{e2, er, 0.0, 0.0}, sh4);
        if (threadIdx.x == 0) {
            double* part4 = PART(0);
            part4[4 * blockIdx.x] = s.a; part4[4 * blockIdx.x + 1] = s.b; part4[4 * blockIdx.x + 2] = 0.0;
            part4[4 * blockIdx.x + 3] = 0.0;
        }
    }
    GRID_SYNC();
    {
        const D4 g = sum_partials4(PART(0), nb, sh4);
        E = g.a + g.b;
    }
    if (tid == 0) {
        sc->E = E;
        sc->E0 = first ? E : E0;
        sc->rz_ref = rz_ref;
        sc->pcg_iters = pcg_total;
        sc->gn_steps = gn_total;
        sc->first = 0;
    }
    PROF(15);
#undef PROF
#undef PART
#undef GRID_SYNC
}

// Version 3 with the rows in REGISTERS (N <= P3_R * resident warps): every warp keeps, for each of its (at most P3_R) rows,
// the column ids / regularisation values / matrix values of P3_LE entries per lane (rows are a few dozen entries long),
// the diagonal, and -- lane c < 3 holding coordinate c -- the PCG vectors r, w, z, s, p, x and the unknown t.  Per PCG
// iteration only the exchanged vector m (one float4 per row) and two doubles per CTA go through memory: one L2 round trip
// for the gathers, one for the partial sums (every WARP sums the CTA partials redundantly in a fixed order, so no CTA-wide
// broadcast is needed), and one grid barrier.
constexpr int P3_R = 2;
constexpr int P3_LE = 2;
constexpr int P3_MAX_LINEAR_ITER = 64;  // longer PCG runs use the textbook recurrences (versions 1 / 2)

DFU_DEV void warp_total4(const double* part4, int nb, int lane, double& a, double& b, double& c) {
    a = b = c = 0.0;
    for (int i = lane; i < nb; i += 32) {
        const double2 x = *reinterpret_cast<const double2*>(part4 + 4 * (size_t) i);
        const double y = part4[4 * (size_t) i + 2];
        a += x.x; b += x.y; c += y;
    }
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
}

__global__ void __launch_bounds__(PTPB, 1) k_solve_persistent3r(Problem pb, Pattern pt, SolveCtl ctl, Scalars* sc, unsigned* bar) {
    constexpr int NWARP = PTPB / 32;
    __shared__ double shw[3 * NWARP];
    __shared__ double tot_sm[3];
    __shared__ unsigned acc_sm[NWARP * 2 * ACC_W];  // per warp: ACC_W low words, ACC_W high words
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, gw = tid >> 5, nw = nthreads >> 5;
    const int nb = gridDim.x, N = pb.N;
    unsigned bar_target = 0;
    unsigned* acc_lo = acc_sm + wib * 2 * ACC_W;
    unsigned* acc_hi = acc_lo + ACC_W;
#define GRID_SYNC() grid_barrier(bar, (unsigned) nb, bar_target)
#define PART(buf) (pb.part + (size_t) (buf) * 2 * MAX_PARTIALS)
    long long t_prev = clock64();
#define PROF(k)                                               \
    do {                                                      \
        if (ctl.prof && tid == 0) {                           \
            const long long t_now = clock64();                \
            ctl.prof[k] += t_now - t_prev;                    \
            t_prev = t_now;                                   \
        }                                                     \
    } while (0)
    // grid barrier whose first warp also sums the per-CTA partials published before it (one reader warp per CTA keeps
    // the 148 x 2368 same-line L2 reads of a fully redundant sum off the critical path); totals broadcast through smem
    auto barrier_totals = [&](const double* part4, double& a, double& b, double& c) {
        __syncthreads();
        if (wib == 0) {
            if (lane == 0) {
                bar_target += (unsigned) nb;
                unsigned seen;
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
                } while (seen < bar_target);
            }
            __syncwarp();
            double x, y, z;
            warp_total4(part4, nb, lane, x, y, z);
            if (lane == 0) {
                tot_sm[0] = x; tot_sm[1] = y; tot_sm[2] = z;
            }
        }
        __syncthreads();
        a = tot_sm[0]; b = tot_sm[1]; c = tot_sm[2];
    };
    // CTA partial of up to three per-lane doubles -> dst[0..2] (fixed order: xor tree per warp, warps ascending)
    auto publish = [&](double a, double b, double c, double* dst) {
        a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
        __syncthreads();  // the previous readers of shw are done
        if (lane == 0) {
            shw[wib] = a; shw[NWARP + wib] = b; shw[2 * NWARP + wib] = c;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double ta = 0.0, tb = 0.0, tc = 0.0;
#pragma unroll
            for (int w = 0; w < NWARP; ++w) {
                ta += shw[w]; tb += shw[NWARP + w]; tc += shw[2 * NWARP + w];
            }
            dst[4 * blockIdx.x] = ta; dst[4 * blockIdx.x + 1] = tb; dst[4 * blockIdx.x + 2] = tc; dst[4 * blockIdx.x + 3] = 0.0;
        }
    };

    // ---- my rows --------------------------------------------------------------------------------------------
    int rn[P3_R], roff[P3_R], rlen[P3_R], rds[P3_R];
    int rc[P3_R][P3_LE];
    float ra[P3_R][P3_LE], rv[P3_R][P3_LE];
    float rD[P3_R], rinv[P3_R];
    double rinvd[P3_R];
    float s_r[P3_R], s_w[P3_R], s_z[P3_R], s_s[P3_R], s_p[P3_R], s_x[P3_R], s_t[P3_R];  // coordinate `lane` (lanes 0..2)
#pragma unroll
    for (int r = 0; r < P3_R; ++r) {
        const int n = gw + r * nw;
        rn[r] = n < N ? n : -1;
        roff[r] = rlen[r] = rds[r] = 0;
        rD[r] = rinv[r] = 0.f;
        rinvd[r] = 0.0;
        s_r[r] = s_w[r] = s_z[r] = s_s[r] = s_p[r] = s_x[r] = s_t[r] = 0.f;
#pragma unroll
        for (int u = 0; u < P3_LE; ++u) {
            rc[r][u] = -1;
            ra[r][u] = rv[r][u] = 0.f;
        }
        if (n < N) {
            roff[r] = pt.rowptr[n];
            rlen[r] = pt.rowlen[n];
            rds[r] = pt.dslot[n];
#pragma unroll
            for (int u = 0; u < P3_LE; ++u) {
                const int j = lane + 32 * u;
                if (j < rlen[r]) {
                    rc[r][u] = pt.col[roff[r] + j];
                    ra[r][u] = pt.areg[roff[r] + j];
                }
            }
            if (lane < 3) pb.t[3 * (size_t) n + lane] = 0.f;  // unknowns := 0 (opt_solver.cpp:192-193)
        }
    }
    GRID_SYNC();

    // row product with the exchanged vector x (float4 per row): registers first, entries beyond 32 * P3_LE from L2
    auto spmv = [&](int r, const float4* x) -> float {
        float ax = 0.f, ay = 0.f, az = 0.f;
        float4 m[P3_LE];
#pragma unroll
        for (int u = 0; u < P3_LE; ++u) m[u] = rc[r][u] >= 0 ? __ldcg(x + rc[r][u]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < P3_LE; ++u) {
            ax = __fmaf_rn(rv[r][u], m[u].x, ax);
            ay = __fmaf_rn(rv[r][u], m[u].y, ay);
            az = __fmaf_rn(rv[r][u], m[u].z, az);
        }
        for (int j = 32 * P3_LE + lane; j < rlen[r]; j += 32) {
            const float v = pt.vals[roff[r] + j];
            const float4 mm = __ldcg(x + pt.col[roff[r] + j]);
            ax = __fmaf_rn(v, mm.x, ax);
            ay = __fmaf_rn(v, mm.y, ay);
            az = __fmaf_rn(v, mm.z, az);
        }
        ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
        return lane == 0 ? ax : (lane == 1 ? ay : az);
    };

    double rz_ref = -1.0, E = 0.0, E0 = 0.0;
    int pcg_total = 0, gn_total = 0;
    bool first = true, stop_all = false;

    for (int outer = 0; outer < ctl.num_iter && !stop_all; ++outer) {
        for (int gn = 0; gn < ctl.nonlinear_iter; ++gn) {
            PROF(0);
            const double e2_local = phase_point_residual(pb, gn == 0, tid, nthreads);
            PROF(1);
            GRID_SYNC();
            PROF(2);
            double rz = 0.0, er = 0.0;
#pragma unroll
            for (int r = 0; r < P3_R; ++r) {
                if (rn[r] < 0) continue;  // uniform over the warp
                const int n = rn[r];
                float ax, ay, az, gx = 0.f, gy = 0.f, gz = 0.f, cnt = 0.f, e2 = 0.f;
                // b_n = sum tw * theta e in 2^40 fixed point (independent of the order of the node's list) ...
                node_gather_data_fixed(pb, n, lane, ax, ay, az);
                PROF(3);
                // ... and, when theta changed, the data part of row n of A (fixed point in shared memory), ACC_W columns per pass
                const bool assemble = gn == 0;
                if (assemble) {
                    const int lo = pb.tptr[n], hi = pb.tptr[n + 1];
                    const bool wide = hi - lo > FIX_MAX_DEG;  // too many contributions for the split words: 64-bit atomics
                    unsigned long long* acc64 = reinterpret_cast<unsigned long long*>(acc_lo);
                    for (int c0 = 0; c0 < rlen[r]; c0 += ACC_W) {
                        for (int j = lane; j < 2 * ACC_W; j += 32) acc_lo[j] = 0u;
                        __syncwarp();
                        for (int e0 = lo + lane; e0 < hi; e0 += 128) {  // 4 entries per lane in flight
                            int v[4];
                            float c[4];
                            uint4 sl[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int e = e0 + 32 * u;
                                const bool ok = e < hi;
                                v[u] = ok ? pb.tv[e] : -1;
                                c[u] = ok ? pb.tw[e] : 0.f;
                                sl[u] = ok ? pt.tslot[e] : make_uint4(0u, 0u, 0u, 0u);
                            }
                            float4 w0[4], w1[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int vv = v[u] >= 0 ? v[u] : 0;
                                c[u] *= v[u] >= 0 ? pb.theta[vv] : 0.f;
                                w0[u] = *(reinterpret_cast<const float4*>(pb.wts) + 2 * (size_t) vv);
                                w1[u] = *(reinterpret_cast<const float4*>(pb.wts) + 2 * (size_t) vv + 1);
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (c[u] == 0.f) continue;
                                const float wk[8] = {w0[u].x, w0[u].y, w0[u].z, w0[u].w, w1[u].x, w1[u].y, w1[u].z, w1[u].w};
                                const unsigned sk[8] = {sl[u].x & 0xffffu, sl[u].x >> 16, sl[u].y & 0xffffu, sl[u].y >> 16,
                                                        sl[u].z & 0xffffu, sl[u].z >> 16, sl[u].w & 0xffffu, sl[u].w >> 16};
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    const unsigned sidx = sk[k] - (unsigned) c0;
                                    if (sidx < (unsigned) ACC_W) {  // 2^40 fixed point as two native 32-bit atomics
                                        const unsigned long long f = (unsigned long long) __float2ll_rn(c[u] * wk[k] * FIX_SCALE);
                                        if (wide) {
                                            atomicAdd(&acc64[sidx], f);
                                        } else {
                                            atomicAdd(&acc_lo[sidx], (unsigned) (f & 0xfffffu));
                                            atomicAdd(&acc_hi[sidx], (unsigned) (f >> 20));
                                        }
                                    }
                                }
                            }
                        }
                        __syncwarp();
                        if (c0 == 0) {
#pragma unroll
                            for (int u = 0; u < P3_LE; ++u)
                                if (rc[r][u] >= 0)
                                    rv[r][u] = ra[r][u] + (wide ? (float) ((double) acc64[lane + 32 * u] * FIX_INV)
                                                                : fix2f(acc_lo[lane + 32 * u], acc_hi[lane + 32 * u]));
                        }
                        for (int j = lane; j < ACC_W && c0 + j < rlen[r]; j += 32)
                            if (c0 + j >= 32 * P3_LE)
                                pt.vals[roff[r] + c0 + j] = pt.areg[roff[r] + c0 + j] +
                                                            (wide ? (float) ((double) acc64[j] * FIX_INV) : fix2f(acc_lo[j], acc_hi[j]));
                        __syncwarp();
                    }
                }
                if (pb.wreg2 > 0.f) {
                    node_gather_reg(pb, n, lane, pb.t, gx, gy, gz, cnt, e2);
                    gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
                    ax -= pb.wreg2 * gx; ay -= pb.wreg2 * gy; az -= pb.wreg2 * gz;
                    e2 = warp_sum(e2);
                }
                if (assemble) {
                    // the diagonal
                    float D;
                    if (rds[r] < 32 * P3_LE) {
                        float pick = rv[r][0];
#pragma unroll
                        for (int u = 1; u < P3_LE; ++u) pick = (rds[r] >> 5) == u ? rv[r][u] : pick;
                        D = __shfl_sync(0xffffffffu, pick, rds[r] & 31);
                    } else {
                        D = pt.vals[roff[r] + rds[r]];
                    }
                    rD[r] = D;
                    rinv[r] = D > 0.f ? 1.f / D : 0.f;
                    rinvd[r] = D > 0.f ? 1.0 / (double) D : 0.0;
                    if (lane == 0) pb.nbuf[3 * (size_t) N + n] = D;
                }
                PROF(4);
                const float b = lane == 0 ? ax : (lane == 1 ? ay : az);
                if (lane < 3) {
                    pb.nbuf[3 * (size_t) n + lane] = b;
                    s_r[r] = b;
                    s_x[r] = 0.f;
                    rz += (double) b * (double) b * rinvd[r];
                    reinterpret_cast<float*>(pt.exch + n)[lane] = b * rinv[r];  // u0 = M^-1 b
                }
                if (lane == 0) er += (double) pb.wreg2 * e2;
            }
            publish(e2_local, rz, er, PART(0));
            PROF(5);
            double ta, tb, tc;
            barrier_totals(PART(0), ta, tb, tc);
            PROF(6);
            const double rz0 = tb;
            E = ta + tc;
            if (first) {
                E0 = E;
                first = false;
            }
            if (rz_ref < 0.0) rz_ref = rz0;
            const bool conv0 = !(rz0 > 0.0) || rz0 <= ctl.tol2 * rz_ref;
            if (ctl.early_out && conv0) {  // converged at this linearisation point
                if (gn == 0 && outer > 0) stop_all = true;
                GRID_SYNC();  // every CTA has read the partials before anyone overwrites them
                break;
            }
            PROF(7);
            if (!conv0) {
#pragma unroll
                for (int r = 0; r < P3_R; ++r) {
                    if (rn[r] < 0) continue;
                    const float w = spmv(r, pt.exch);  // w0 = A u0
                    s_w[r] = lane < 3 ? w : 0.f;
                    s_z[r] = s_s[r] = s_p[r] = 0.f;
                }
                double gamma_prev = 0.0, alpha_prev = 0.0;
                PROF(8);
                for (int it = 0; it < ctl.linear_iter; ++it) {
                    const int buf = (it + 1) & 1;
                    float4* ex = pt.exch + (size_t) buf * N;
                    float2* partf = reinterpret_cast<float2*>(PART(buf));  // per-CTA (g, d) as floats: alpha, beta are floats anyway
                    double g = 0.0, d = 0.0;
#pragma unroll
                    for (int r = 0; r < P3_R; ++r) {
                        if (rn[r] < 0) continue;
                        if (lane < 3) {
                            g += (double) s_r[r] * (double) s_r[r] * rinvd[r];
                            d += (double) s_w[r] * (double) s_r[r] * rinvd[r];
                            reinterpret_cast<float*>(ex + rn[r])[lane] = s_w[r] * rinv[r];  // m = M^-1 w
                        }
                    }
                    g = warp_sum(g); d = warp_sum(d);
                    __syncthreads();
                    if (lane == 0) {
                        shw[wib] = g; shw[NWARP + wib] = d;
                    }
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        double tg = 0.0, td = 0.0;
#pragma unroll
                        for (int w = 0; w < NWARP; ++w) {
                            tg += shw[w]; td += shw[NWARP + w];
                        }
                        partf[blockIdx.x] = make_float2((float) tg, (float) td);
                    }
                    PROF(9);
                    GRID_SYNC();
                    PROF(10);
                    // after the barrier: the partial sums (first warp, loads issued first) and the row products n = A m,
                    // which do not depend on this iteration's scalars, share one L2 round trip
                    float2 pf[5];
                    if (wib == 0) {
#pragma unroll
                        for (int u = 0; u < 5; ++u) {
                            const int i = lane + 32 * u;
                            pf[u] = i < nb ? __ldcg(partf + i) : make_float2(0.f, 0.f);
                        }
                    }
                    const bool last = it + 1 >= ctl.linear_iter;
                    float nv[P3_R];
#pragma unroll
                    for (int r = 0; r < P3_R; ++r) nv[r] = (rn[r] >= 0 && !last) ? spmv(r, ex) : 0.f;
                    if (wib == 0) {
                        double a = 0.0, b2 = 0.0;
#pragma unroll
                        for (int u = 0; u < 5; ++u) {
                            a += (double) pf[u].x; b2 += (double) pf[u].y;
                        }
                        for (int i = lane + 160; i < nb; i += 32) {  // more than 160 CTAs: the rest
                            const float2 x = __ldcg(partf + i);
                            a += (double) x.x; b2 += (double) x.y;
                        }
                        a = warp_sum(a); b2 = warp_sum(b2);
                        if (lane == 0) {
                            tot_sm[0] = a; tot_sm[1] = b2;
                        }
                    }
                    __syncthreads();
                    const double gamma = tot_sm[0], delta = tot_sm[1];
                    PROF(11);
                    if (!(gamma > 0.0) || (it > 0 && gamma <= ctl.tol2 * rz_ref)) break;
                    const double beta = it > 0 ? gamma / gamma_prev : 0.0;
                    const double denom = it > 0 ? delta - beta * gamma / alpha_prev : delta;
                    if (!(denom > 0.0)) break;
                    const double alpha = gamma / denom;
                    const float af = (float) alpha, bf = (float) beta;
#pragma unroll
                    for (int r = 0; r < P3_R; ++r) {
                        if (rn[r] < 0) continue;
                        if (lane < 3) {
                            s_z[r] = __fmaf_rn(bf, s_z[r], nv[r]);
                            s_s[r] = __fmaf_rn(bf, s_s[r], s_w[r]);
                            s_p[r] = __fmaf_rn(bf, s_p[r], s_r[r] * rinv[r]);
                            s_x[r] = __fmaf_rn(af, s_p[r], s_x[r]);
                            s_r[r] = __fmaf_rn(-af, s_s[r], s_r[r]);
                            s_w[r] = __fmaf_rn(-af, s_z[r], s_w[r]);
                        }
                    }
                    PROF(12);
                    ++pcg_total;
                    gamma_prev = gamma;
                    alpha_prev = alpha;
                }
            }
#pragma unroll
            for (int r = 0; r < P3_R; ++r)
                if (rn[r] >= 0 && lane < 3) {
                    s_t[r] += s_x[r];
                    pb.t[3 * (size_t) rn[r] + lane] = s_t[r];
                }
            ++gn_total;
            PROF(13);
            GRID_SYNC();
            PROF(14);
        }
    }
    // ---- final energy at the solution, Tukey weights of the last outer iteration -------------------------
    {
        const double e2 = phase_point_residual(pb, first, tid, nthreads);  // no GN step ran: weights at t = 0
        double er = 0.0;
        if (pb.wreg2 > 0.f) {
#pragma unroll
            for (int r = 0; r < P3_R; ++r) {
                if (rn[r] < 0) continue;
                float gx, gy, gz, cnt, r2;
                node_gather_reg(pb, rn[r], lane, pb.t, gx, gy, gz, cnt, r2);
                r2 = warp_sum(r2);
                if (lane == 0) er += (double) pb.wreg2 * r2;
            }
        }
        publish(e2, er, 0.0, PART(0));
    }
    {
        double ta, tb, tc;
        barrier_totals(PART(0), ta, tb, tc);
        E = ta + tb;
    }
    if (tid == 0) {
        sc->E = E;
        sc->E0 = first ? E : E0;
        sc->rz_ref = rz_ref;
        sc->pcg_iters = pcg_total;
        sc->gn_steps = gn_total;
        sc->first = 0;
    }
    PROF(15);
#undef PROF
#undef PART
#undef GRID_SYNC
}

// Sparsity pattern of A for one frame, one warp per row: a bitmap over the nodes (shared memory) collects the
// diagonal, the regularisation edges in both directions and the 8 neighbours of every point that references the
// node; its set bits in ascending order are the columns.  Rows are allocated with one atomic per row, so their
// order in memory is arbitrary (nothing depends on it).  Also emitted: the regularisation values w_reg^2 L, the slot
// of the diagonal, and for every transposed-graph entry the slots of its point's 8 neighbours.
__global__ void __launch_bounds__(128) k_pattern(Problem pb, int NW, int* __restrict__ cursor, int* __restrict__ rowptr,
                                                 int* __restrict__ rowlen, int* __restrict__ dslot, int32_t* __restrict__ col,
                                                 float* __restrict__ areg, uint4* __restrict__ tslot) {
    extern __shared__ unsigned pat_sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned* bm = pat_sm + (size_t) wib * 2 * NW;
    unsigned* pf = bm + NW;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    auto slot_of = [&](int m) -> unsigned { return pf[m >> 5] + __popc(bm[m >> 5] & ((1u << (m & 31)) - 1u)); };
    for (int a = gw; a < pb.N; a += nw) {
        for (int w = lane; w < NW; w += 32) bm[w] = 0u;
        __syncwarp();
        const int lo = pb.tptr[a], hi = pb.tptr[a + 1];
        const int rlo = pb.rin_ptr[a], rhi = pb.rin_ptr[a + 1];
        if (lane == 0) atomicOr(&bm[a >> 5], 1u << (a & 31));
        if (lane < 8) {
            const int m = pb.nnbr[(size_t) a * 8 + lane];
            atomicOr(&bm[m >> 5], 1u << (m & 31));
        }
        for (int j = rlo + lane; j < rhi; j += 32) {
            const int m = pb.rin[j];
            atomicOr(&bm[m >> 5], 1u << (m & 31));
        }
        for (int e = lo + lane; e < hi; e += 32) {
            int nbk[8];
            float wk[8];
            load8(pb.nbr, pb.wts, pb.tv[e], nbk, wk);
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicOr(&bm[nbk[k] >> 5], 1u << (nbk[k] & 31));
        }
        __syncwarp();
        // exclusive prefix of the word popcounts
        int base = 0;
        for (int w0 = 0; w0 < NW; w0 += 32) {
            const int w = w0 + lane;
            const int c = w < NW ? __popc(bm[w]) : 0;
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (w < NW) pf[w] = (unsigned) (base + inc - c);
            base += __shfl_sync(0xffffffffu, inc, 31);
        }
        const int len = base;
        int off = 0;
        if (lane == 0) off = atomicAdd(cursor, len);
        off = __shfl_sync(0xffffffffu, off, 0);
        __syncwarp();
        for (int w = lane; w < NW; w += 32) {
            unsigned bits = bm[w];
            int j = (int) pf[w];
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                col[off + j] = 32 * w + b;
                areg[off + j] = 0.f;
                ++j;
            }
        }
        __syncwarp();
        // regularisation part: -w_reg^2 per edge (either direction, duplicates add up), + w_reg^2 * edges on the diagonal.
        // All addends of one entry are equal, so the order of the atomics cannot change the sum.
        const int ds = (int) slot_of(a);
        float cnt = 0.f;
        if (pb.wreg2 > 0.f)
            for (int j = lane; j < 8 + (rhi - rlo); j += 32) {
                const int m = j < 8 ? pb.nnbr[(size_t) a * 8 + j] : pb.rin[rlo + j - 8];
                if (m == a) continue;
                atomicAdd(&areg[off + slot_of(m)], -pb.wreg2);
                cnt += 1.f;
            }
        cnt = warp_sum(cnt);
        __syncwarp();
        if (lane == 0) {
            rowptr[a] = off;
            rowlen[a] = len;
            dslot[a] = ds;
            areg[off + ds] = pb.wreg2 * cnt;
        }
        for (int e = lo + lane; e < hi; e += 32) {
            int nbk[8];
            float wk[8];
            load8(pb.nbr, pb.wts, pb.tv[e], nbk, wk);
            unsigned sk[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) sk[k] = slot_of(nbk[k]);
            tslot[e] = make_uint4(sk[0] | (sk[1] << 16), sk[2] | (sk[3] << 16), sk[4] | (sk[5] << 16), sk[6] | (sk[7] << 16));
        }
        __syncwarp();
    }
}

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

}  // namespace

struct dfu_solver {
    dfu_warpfield* wf = nullptr;
    int device = 0;
    dfu_solver_params prm{};
    dfu_allreduce_fn allreduce = nullptr;
    void* allreduce_ctx = nullptr;
    int N = 0, P = 0;
    size_t capP = 0, capN = 0;
    // per point
    int32_t* nbr = nullptr;
    float *wts = nullptr, *dvec = nullptr, *theta = nullptr;
    float4* s4 = nullptr;
    int32_t *tent = nullptr, *tv = nullptr;
    float* tw = nullptr;
    // per node
    int* tptr = nullptr;
    int* tmp = nullptr;  // [deg N | cursor N]
    int32_t* nnbr = nullptr;
    int* rin_ptr = nullptr;
    int32_t* rin = nullptr;
    float* vec = nullptr;   // t, dl, r, z, p, q : 6 * 3N
    float* nbuf = nullptr;  // 4N + 4
    double* part = nullptr;
    Scalars* sc = nullptr;
    Scalars* sc_host = nullptr;  // pinned
    unsigned* bar = nullptr;     // grid-barrier counter of the persistent kernel
    uint64_t reg_epoch = 0;      // node-position epoch the regularisation graph was built for (0: never)
    bool problem_ready = false;
    int coop_blocks = 0;         // co-resident CTAs for the persistent kernel (0: not available)
    int coop_blocks2 = 0;        // same for version 2 of the kernel
    int coop_blocks3 = 0;        // same for version 3
    int coop_blocks3r = 0;       // same for version 3 with the rows in registers
    int last_kernel = 0;         // which persistent kernel the last solve used (1 / 2 / 3; 0: multi-kernel path)
    // explicit normal matrix (version 3): pattern per frame, values per re-weighting
    int *rowptr = nullptr, *rowlen = nullptr, *dslot = nullptr, *pat_cursor = nullptr;
    int32_t* col = nullptr;
    float *areg = nullptr, *vals = nullptr;
    uint4* tslot = nullptr;
    float4 *exch = nullptr, *st = nullptr;
    unsigned long long *xw = nullptr, *pw = nullptr;
    size_t cap_nnz = 0, cap_slots = 0, cap_rows = 0;
    bool pattern_ready = false;
    bool lists_sorted = true;    // transposed lists ordered by point id (needed by the float-order-dependent paths)
    int gn_steps_host = 0;
};

namespace {

void free_point_arrays(dfu_solver* s) {
    cudaFree(s->nbr); cudaFree(s->wts); cudaFree(s->dvec); cudaFree(s->theta); cudaFree(s->s4);
    cudaFree(s->tent); cudaFree(s->tv); cudaFree(s->tw);
    s->nbr = s->tent = s->tv = nullptr;
    s->wts = s->dvec = s->theta = s->tw = nullptr;
    s->s4 = nullptr;
    s->capP = 0;
}
void free_pattern_arrays(dfu_solver* s) {
    cudaFree(s->rowptr); cudaFree(s->rowlen); cudaFree(s->dslot); cudaFree(s->col); cudaFree(s->areg); cudaFree(s->vals);
    cudaFree(s->tslot); cudaFree(s->exch); cudaFree(s->st); cudaFree(s->xw); cudaFree(s->pw);
    s->xw = s->pw = nullptr;
    s->rowptr = s->rowlen = s->dslot = nullptr;
    s->col = nullptr;
    s->areg = s->vals = nullptr;
    s->tslot = nullptr;
    s->exch = s->st = nullptr;
    s->cap_nnz = s->cap_slots = s->cap_rows = 0;
    s->pattern_ready = false;
}
void free_node_arrays(dfu_solver* s) {
    cudaFree(s->tptr); cudaFree(s->tmp); cudaFree(s->nnbr); cudaFree(s->rin_ptr); cudaFree(s->rin); cudaFree(s->vec);
    cudaFree(s->nbuf);
    s->tptr = s->tmp = s->rin_ptr = nullptr;
    s->nnbr = s->rin = nullptr;
    s->vec = s->nbuf = nullptr;
    s->capN = 0;
    s->reg_epoch = 0;
}

int read_scalars(dfu_solver* s, cudaStream_t st) {
    DFU_CUDA_OK(cudaMemcpyAsync(s->sc_host, s->sc, sizeof(Scalars), cudaMemcpyDeviceToHost, st));
    DFU_CUDA_OK(cudaStreamSynchronize(st));
    return DFU_OK;
}

Problem make_problem(const dfu_solver* s) {
    const size_t n3 = 3 * (size_t) s->N;
    Problem pb{};
    pb.N = s->N; pb.P = s->P;
    pb.nbr = s->nbr; pb.wts = s->wts; pb.dvec = s->dvec; pb.theta = s->theta; pb.s4 = s->s4;
    pb.tptr = s->tptr; pb.tv = s->tv; pb.tw = s->tw;
    pb.nnbr = s->nnbr; pb.rin_ptr = s->rin_ptr; pb.rin = s->rin;
    pb.wreg2 = s->prm.lambda / ((float) s->N * 8.f);  // w_reg^2 (opt_solver.cpp:30)
    pb.t = s->vec; pb.dl = s->vec + n3; pb.r = s->vec + 2 * n3; pb.z = s->vec + 3 * n3; pb.p = s->vec + 4 * n3;
    pb.q = s->vec + 5 * n3;
    pb.nbuf = s->nbuf;
    pb.part = s->part;
    pb.tukey_offset = s->prm.tukey_offset;
    pb.psi_data = s->prm.psi_data;
    return pb;
}

// multi-kernel path: residuals + block assembly (+ all-reduce) + regularisation + PCG initialisation
int assemble_mk(dfu_solver* s, const Problem& pb, bool update_tukey, cudaStream_t st) {
    const int N = s->N, P = s->P;
    const int nblk_p = P > 0 ? min(div_up(P, TPB), MAX_PARTIALS) : 0;
    const int nblk_w = min(div_up((long) N * 32, TPB), MAX_PARTIALS);
    if (P > 0) {
        k_point_residual<<<nblk_p, TPB, 0, st>>>(pb, update_tukey ? 1 : 0);
        DFU_LAUNCH_OK();
    }
    k_node_assemble_data<<<nblk_w, TPB, 0, st>>>(pb, nblk_p);
    DFU_LAUNCH_OK();
    if (s->allreduce) {
        int rc = s->allreduce(s->nbuf, 4 * (size_t) N + 4, s->allreduce_ctx, (dfu_stream) st);
        DFU_REQUIRE(rc == 0, DFU_ERR_CUDA, "all-reduce hook failed");
    }
    k_node_reg_init<<<nblk_w, TPB, 0, st>>>(pb);
    DFU_LAUNCH_OK();
    const double tol2 = (double) s->prm.pcg_tol * (double) s->prm.pcg_tol;
    k_init_scalars<<<1, 32, 0, st>>>(pb, s->sc, nblk_w, tol2);
    DFU_LAUNCH_OK();
    return DFU_OK;
}

int solve_multi_kernel(dfu_solver* s, cudaStream_t st) {
    const int N = s->N, P = s->P;
    const dfu_solver_params& prm = s->prm;
    const Problem pb = make_problem(s);
    const int nblk_n = div_up(N, TPB);
    const int nblk_p = P > 0 ? min(div_up(P, TPB), MAX_PARTIALS) : 0;
    const int nblk_w = min(div_up((long) N * 32, TPB), MAX_PARTIALS);
    const double tol2 = (double) prm.pcg_tol * (double) prm.pcg_tol;
    const bool host_checks = prm.early_out != 0;

    Scalars init{};
    init.rz_ref = -1.0;
    init.first = 1;
    init.done_it = INT_MAX;
    *s->sc_host = init;
    DFU_CUDA_OK(cudaMemcpyAsync(s->sc, s->sc_host, sizeof(Scalars), cudaMemcpyHostToDevice, st));
    DFU_CUDA_OK(cudaMemsetAsync(pb.t, 0, 3 * (size_t) N * sizeof(float), st));
    if (host_checks) DFU_CUDA_OK(cudaStreamSynchronize(st));  // sc_host is re-used for read-backs below

    bool stop_all = false;
    s->gn_steps_host = 0;
    for (int outer = 0; outer < prm.num_iter && !stop_all; ++outer) {
        for (int gn = 0; gn < prm.nonlinear_iter; ++gn) {
            int rc = assemble_mk(s, pb, gn == 0, st);  // preNonlinearSolve re-weights once per outer iteration
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
                    k_point_apply<<<nblk_p, TPB, 0, st>>>(pb, s->sc, it);
                    DFU_LAUNCH_OK();
                }
                k_node_apply_data<<<nblk_w, TPB, 0, st>>>(pb, s->sc, it);
                DFU_LAUNCH_OK();
                if (s->allreduce) {
                    rc = s->allreduce(pb.q, 3 * (size_t) N, s->allreduce_ctx, (dfu_stream) st);
                    DFU_REQUIRE(rc == 0, DFU_ERR_CUDA, "all-reduce hook failed");
                }
                k_node_apply_reg_dot<<<nblk_w, TPB, 0, st>>>(pb, s->sc, it);
                DFU_LAUNCH_OK();
                k_pcg_update<<<nblk_n, TPB, 0, st>>>(pb, s->sc, it, nblk_w);
                DFU_LAUNCH_OK();
                k_pcg_direction<<<nblk_n, TPB, 0, st>>>(pb, s->sc, it, nblk_n, nblk_w, tol2);
                DFU_LAUNCH_OK();
                if (host_checks && (it & 7) == 7) {
                    rc = read_scalars(s, st);
                    if (rc != DFU_OK) return rc;
                    if (s->sc_host->done_it <= it + 1) break;
                }
            }
            k_axpy<<<div_up(3L * N, TPB), TPB, 0, st>>>(pb.t, pb.dl, 3 * N);
            DFU_LAUNCH_OK();
            s->gn_steps_host += 1;
        }
    }
    return assemble_mk(s, pb, s->gn_steps_host == 0, st);  // final energy at the solution (tukey weights at t = 0 if no step ran)
}

int solve_persistent(dfu_solver* s, cudaStream_t st) {
    Problem pb = make_problem(s);
    SolveCtl ctl{s->prm.num_iter, s->prm.nonlinear_iter, s->prm.linear_iter, s->prm.early_out,
                 (double) s->prm.pcg_tol * (double) s->prm.pcg_tol, nullptr};
    Scalars* sc = s->sc;
    unsigned* bar = s->bar;
    DFU_CUDA_OK(cudaMemsetAsync(bar, 0, sizeof(unsigned), st));
    static long long* prof_dev = nullptr;
    const bool profile = getenv("DFU_SOLVER_PROFILE") != nullptr;
    if (profile) {
        if (!prof_dev) DFU_CUDA_OK(cudaMalloc(&prof_dev, 16 * sizeof(long long)));
        DFU_CUDA_OK(cudaMemsetAsync(prof_dev, 0, 16 * sizeof(long long), st));
        ctl.prof = prof_dev;
    }
    // DFU_SOLVER_PATH = p1 | p2 | p3 forces a version of the persistent kernel (tests cover all of them).
    //   3: explicit normal matrix + pipelined PCG (needs the pattern built by init_problem)
    //   2: matrix-free, graph in registers (needs every node to fit a register slot: N <= 2 * resident warps)
    //   1: matrix-free, everything from L2
    const char* force = getenv("DFU_SOLVER_PATH");
    const bool forced = force && force[0] == 'p' && force[1] >= '1' && force[1] <= '3';
    // Default: version 3 for the short fixed-budget solves of the frame loop.  Its pipelined recurrences lose attainable
    // accuracy in long runs (Ghysels & Vanroose 2014, sec. 4), so tolerance-driven solves with a large iteration budget
    // take the textbook PCG of version 2 / 1.
    const int want = forced ? force[1] - '0' : (s->prm.linear_iter <= P3_MAX_LINEAR_ITER ? 3 : 2);
    int ver = want;
    if (ver == 3 && !(s->pattern_ready && s->coop_blocks3 > 0 && s->N <= 32 * s->coop_blocks3 * (PTPB / 32))) ver = 2;  // (lane-per-row blocks)
    if (ver == 2 && (s->coop_blocks2 == 0 || s->N > P2_NPW * s->coop_blocks2 * (PTPB / 32))) ver = 1;
    if (ver == 3) {
        Pattern pt{s->rowptr, s->rowlen, s->dslot, s->col, s->areg, s->vals, s->tslot, s->exch, s->st, s->xw, s->pw};
        void* args[] = {&pb, &pt, &ctl, &sc, &bar};
        // rows in registers when every node fits a register slot; DFU_SOLVER_PATH=p3g forces the generic kernel
        const bool reg = s->coop_blocks3r > 0 && s->N <= P3_R * s->coop_blocks3r * (PTPB / 32) && !(force && force[1] == '3' && force[2] == 'g');
        if (reg) {  // tags restart at 1 every launch: stale words of earlier launches must not look fresh
            DFU_CUDA_OK(cudaMemsetAsync(s->xw, 0, 2 * 3 * (size_t) s->N * sizeof(unsigned long long), st));
            DFU_CUDA_OK(cudaMemsetAsync(s->pw, 0, 2 * 2 * (size_t) MAX_PARTIALS * sizeof(unsigned long long), st));
            DFU_CUDA_OK(cudaMemsetAsync(&s->sc->spin_fail, 0, sizeof(int), st));
        }
        if (reg)
            DFU_CUDA_OK(cudaLaunchCooperativeKernel((void*) k_solve_persistent3r, dim3(s->coop_blocks3r), dim3(PTPB), args, 0, st));
        else
            DFU_CUDA_OK(cudaLaunchCooperativeKernel((void*) k_solve_persistent3, dim3(s->coop_blocks3), dim3(PTPB), args, 0, st));
        if (getenv("DFU_DEBUG")) fprintf(stderr, "[dfu] v3 %s\n", reg ? "rows in registers" : "generic");
    } else {
        void* args[] = {&pb, &ctl, &sc, &bar};
        if (ver == 1)
            DFU_CUDA_OK(cudaLaunchCooperativeKernel((void*) k_solve_persistent, dim3(s->coop_blocks), dim3(PTPB), args, 0, st));
        else
            DFU_CUDA_OK(cudaLaunchCooperativeKernel((void*) k_solve_persistent2, dim3(s->coop_blocks2), dim3(PTPB), args, 0, st));
    }
    s->last_kernel = ver;
    ++g_dfu_launches;
    if (profile && ver == 3) {  // debugging aid: synchronises
        long long h[16];
        DFU_CUDA_OK(cudaMemcpyAsync(h, prof_dev, sizeof(h), cudaMemcpyDeviceToHost, st));
        DFU_CUDA_OK(cudaStreamSynchronize(st));
        static const char* names[16] = {"misc", "residual", "bar", "gather_b", "assemble", "row_init", "bar", "sum_part", "spmv_w0",
                                        "dots+m", "bar", "sum_part", "spmv+update", "t+=x", "bar", "final"};
        fprintf(stderr, "[dfu] v3 cycles of CTA 0:");
        for (int i = 0; i < 16; ++i) fprintf(stderr, " %s=%lld", names[i], h[i]);
        fprintf(stderr, "\n");
    }
    s->gn_steps_host = -1;  // read from the device scalars
    if (getenv("DFU_DEBUG")) fprintf(stderr, "[dfu] persistent solver kernel v%d\n", s->last_kernel);
    return DFU_OK;
}

// sparsity pattern of the explicit normal matrix for this frame's graphs (version 3 of the persistent kernel)
// can this problem run version 3 of the persistent kernel (explicit normal matrix)?
bool pattern_eligible(const dfu_solver* s) {
    const char* force = getenv("DFU_SOLVER_PATH");
    if (force && (force[0] == 'm' || (force[0] == 'p' && (force[1] == '1' || force[1] == '2')))) return false;  // another path was asked for
    const bool forced3 = force && force[0] == 'p' && force[1] == '3';
    if (!forced3 && s->prm.linear_iter > P3_MAX_LINEAR_ITER) return false;
    return !(s->coop_blocks3 == 0 || s->allreduce != nullptr || s->N > 65535 || (long) s->P >= (1L << 19));
}

int build_pattern(dfu_solver* s, cudaStream_t st) {
    s->pattern_ready = false;
    const int N = s->N, P = s->P;
    if (!pattern_eligible(s)) return DFU_OK;
    // hard upper bound of the non-zeros: every (node, point) pair contributes at most 8 columns, plus 16 edges + diagonal
    const size_t nnz_cap = std::min<size_t>((size_t) 64 * P + (size_t) 17 * N, (size_t) N * N);
    const size_t slots = (size_t) 8 * P;
    if (nnz_cap > s->cap_nnz) {
        cudaFree(s->col); cudaFree(s->areg); cudaFree(s->vals);
        s->col = nullptr; s->areg = s->vals = nullptr; s->cap_nnz = 0;
        DFU_CUDA_OK(cudaMalloc(&s->col, nnz_cap * sizeof(int32_t)));
        DFU_CUDA_OK(cudaMalloc(&s->areg, nnz_cap * sizeof(float)));
        DFU_CUDA_OK(cudaMalloc(&s->vals, nnz_cap * sizeof(float)));
        s->cap_nnz = nnz_cap;
    }
    if (slots > s->cap_slots) {
        cudaFree(s->tslot);
        s->tslot = nullptr; s->cap_slots = 0;
        DFU_CUDA_OK(cudaMalloc(&s->tslot, std::max<size_t>(slots, 1) * sizeof(uint4)));
        s->cap_slots = slots;
    }
    if ((size_t) N > s->cap_rows) {
        cudaFree(s->rowptr); cudaFree(s->rowlen); cudaFree(s->dslot); cudaFree(s->exch); cudaFree(s->st); cudaFree(s->xw);
        s->rowptr = s->rowlen = s->dslot = nullptr; s->exch = s->st = nullptr; s->xw = nullptr; s->cap_rows = 0;
        DFU_CUDA_OK(cudaMalloc(&s->rowptr, (size_t) N * sizeof(int)));
        DFU_CUDA_OK(cudaMalloc(&s->rowlen, (size_t) N * sizeof(int)));
        DFU_CUDA_OK(cudaMalloc(&s->dslot, (size_t) N * sizeof(int)));
        DFU_CUDA_OK(cudaMalloc(&s->exch, 2 * (size_t) N * sizeof(float4)));
        DFU_CUDA_OK(cudaMalloc(&s->st, 6 * (size_t) N * sizeof(float4)));
        DFU_CUDA_OK(cudaMalloc(&s->xw, 2 * 3 * (size_t) N * sizeof(unsigned long long)));
        s->cap_rows = (size_t) N;
    }
    if (!s->pat_cursor) DFU_CUDA_OK(cudaMalloc(&s->pat_cursor, sizeof(int)));
    if (!s->pw) DFU_CUDA_OK(cudaMalloc(&s->pw, 2 * 2 * (size_t) MAX_PARTIALS * sizeof(unsigned long long)));
    DFU_CUDA_OK(cudaMemsetAsync(s->pat_cursor, 0, sizeof(int), st));
    const int NW = (N + 31) / 32;
    const size_t smem = (size_t) 4 * 2 * NW * sizeof(unsigned);
    static bool attr_set = false;
    if (smem > 48 * 1024 && !attr_set) {
        DFU_CUDA_OK(cudaFuncSetAttribute(k_pattern, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 2 * 2048 * (int) sizeof(unsigned)));
        attr_set = true;
    }
    const Problem pb = make_problem(s);
    k_pattern<<<div_up(N, 4), 128, smem, st>>>(pb, NW, s->pat_cursor, s->rowptr, s->rowlen, s->dslot, s->col, s->areg, s->tslot);
    DFU_LAUNCH_OK();
    s->pattern_ready = true;
    if (getenv("DFU_DEBUG")) {  // debugging aid: synchronises
        std::vector<int> len((size_t) N);
        DFU_CUDA_OK(cudaMemcpyAsync(len.data(), s->rowlen, (size_t) N * sizeof(int), cudaMemcpyDeviceToHost, st));
        DFU_CUDA_OK(cudaStreamSynchronize(st));
        long tot = 0;
        int mx = 0, over64 = 0, over128 = 0;
        for (int v : len) {
            tot += v;
            mx = std::max(mx, v);
            over64 += v > 64;
            over128 += v > 128;
        }
        fprintf(stderr, "[dfu] normal-matrix pattern: N %d nnz %ld mean row %.1f max row %d rows>64 %d rows>128 %d\n", N, tot,
                (double) tot / N, mx, over64, over128);
    }
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
    s->device = wf->device;
    s->prm = *prm;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(wf->device);
    cudaError_t e1 = cudaMalloc(&s->sc, sizeof(Scalars));
    cudaError_t e2 = cudaMallocHost(&s->sc_host, sizeof(Scalars));
    cudaError_t e3 = cudaMalloc(&s->part, 4 * MAX_PARTIALS * sizeof(double));
    if (e3 == cudaSuccess) e3 = cudaMalloc(&s->bar, 64);
    int coop = 0, sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, wf->device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, wf->device);
    if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_persistent, PTPB, 0) == cudaSuccess && per_sm >= 1)
        s->coop_blocks = min(sms, MAX_PARTIALS);  // one CTA per SM
    if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_persistent2, PTPB, 0) == cudaSuccess && per_sm >= 1)
        s->coop_blocks2 = min(sms, MAX_PARTIALS);
    if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_persistent3, PTPB, 0) == cudaSuccess && per_sm >= 1)
        s->coop_blocks3 = min(sms, MAX_PARTIALS);
    if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_persistent3r, PTPB, 0) == cudaSuccess && per_sm >= 1)
        s->coop_blocks3r = min(sms, MAX_PARTIALS);
    (void) cudaGetLastError();
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
    cudaSetDevice(s->device);  // not s->wf->device: the warp field may already be gone
    free_point_arrays(s);
    free_node_arrays(s);
    free_pattern_arrays(s);
    cudaFree(s->pat_cursor);
    cudaFree(s->sc);
    cudaFreeHost(s->sc_host);
    cudaFree(s->part);
    cudaFree(s->bar);
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
    DFU_REQUIRE((long) P * 8 < 0x7fffffffL, DFU_ERR_INVALID, "too many points");
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != wf->device) DFU_CUDA_OK(cudaSetDevice(wf->device));
    (void) cudaGetLastError();  // drop stale errors of other libraries
    cudaStream_t st = as_stream(stream);
    const int N = wf->N;
    if ((size_t) P > s->capP) {
        free_point_arrays(s);
        const size_t cap = (size_t) P;
        DFU_CUDA_OK(cudaMalloc(&s->nbr, cap * 8 * sizeof(int32_t)));
        DFU_CUDA_OK(cudaMalloc(&s->wts, cap * 8 * sizeof(float)));
        DFU_CUDA_OK(cudaMalloc(&s->dvec, cap * 3 * sizeof(float)));
        DFU_CUDA_OK(cudaMalloc(&s->theta, cap * sizeof(float)));
        DFU_CUDA_OK(cudaMalloc(&s->s4, cap * sizeof(float4)));
        DFU_CUDA_OK(cudaMalloc(&s->tent, cap * 8 * sizeof(int32_t)));
        DFU_CUDA_OK(cudaMalloc(&s->tv, cap * 8 * sizeof(int32_t)));
        DFU_CUDA_OK(cudaMalloc(&s->tw, cap * 8 * sizeof(float)));
        s->capP = cap;
    }
    if ((size_t) N > s->capN) {
        free_node_arrays(s);
        const size_t cap = (size_t) N;
        DFU_CUDA_OK(cudaMalloc(&s->tptr, (cap + 1) * sizeof(int)));
        DFU_CUDA_OK(cudaMalloc(&s->tmp, 2 * cap * sizeof(int)));
        DFU_CUDA_OK(cudaMalloc(&s->nnbr, cap * 8 * sizeof(int32_t)));
        DFU_CUDA_OK(cudaMalloc(&s->rin_ptr, (cap + 1) * sizeof(int)));
        DFU_CUDA_OK(cudaMalloc(&s->rin, cap * 8 * sizeof(int32_t)));
        DFU_CUDA_OK(cudaMalloc(&s->vec, 18 * cap * sizeof(float)));
        DFU_CUDA_OK(cudaMalloc(&s->nbuf, (4 * cap + 4) * sizeof(float)));
        s->capN = cap;
    }
    s->N = N;
    s->P = P;
    int rc = DFU_OK;
    // regularisation graph (opt_solver.cpp:74-105) and its transpose: depend on node POSITIONS only, so they are
    // rebuilt exactly when the reference would see a different KD-tree (Warpfield::init / update)
    if (s->reg_epoch != wf->node_epoch) {
        rc = dfu_wf_build_node_graph(wf, s->nnbr, st);
        if (rc != DFU_OK) return rc;
        const long ne = (long) N * 8;
        DFU_CUDA_OK(cudaMemsetAsync(s->tmp, 0, 2 * (size_t) N * sizeof(int), st));
        k_count<<<div_up(ne, TPB), TPB, 0, st>>>(s->nnbr, ne, s->tmp);
        DFU_LAUNCH_OK();
        k_scan<<<1, 1024, 0, st>>>(s->tmp, N, s->rin_ptr);
        DFU_LAUNCH_OK();
        k_fill<<<div_up(ne, TPB), TPB, 0, st>>>(s->nnbr, ne, s->rin_ptr, s->tmp + N, s->rin, 3);
        DFU_LAUNCH_OK();
        k_sort_small<<<div_up(N, TPB), TPB, 0, st>>>(s->rin_ptr, N, s->rin);
        DFU_LAUNCH_OK();
        s->reg_epoch = wf->node_epoch;
    }
    // data graph (opt_solver.cpp:56-72) with the per-edge weights (the kNN kernel also counts the references per node),
    // and its transpose: per node the (point, weight) pairs -- in arrival order when only order-independent consumers
    // will read them (versions 3 / 3r), sorted by point otherwise
    const long ne = (long) P * 8;
    DFU_CUDA_OK(cudaMemsetAsync(s->tmp, 0, 2 * (size_t) N * sizeof(int), st));
    if (P > 0) {
        rc = dfu_wf_build_data_graph(wf, canon_v, live_v, P, s->nbr, s->wts, s->dvec, s->tmp, st);
        if (rc != DFU_OK) return rc;
    }
    k_scan<<<1, 1024, 0, st>>>(s->tmp, N, s->tptr);
    DFU_LAUNCH_OK();
    s->lists_sorted = !pattern_eligible(s);
    if (P > 0) {
        if (s->lists_sorted) {
            k_fill<<<div_up(ne, TPB), TPB, 0, st>>>(s->nbr, ne, s->tptr, s->tmp + N, s->tent, 0);
            DFU_LAUNCH_OK();
            k_sort_emit<<<min(div_up((long) N * 32, TPB), 65535), TPB, 0, st>>>(s->tptr, N, s->tent, s->wts, s->tv, s->tw);
            DFU_LAUNCH_OK();
        } else {
            k_fill_emit<<<div_up(ne, TPB), TPB, 0, st>>>(s->nbr, ne, s->tptr, s->tmp + N, s->wts, s->tv, s->tw);
            DFU_LAUNCH_OK();
        }
    }
    DFU_CUDA_OK(cudaMemsetAsync(s->vec, 0, 18 * (size_t) N * sizeof(float), st));  // unknowns := 0 (opt_solver.cpp:192-193)
    rc = build_pattern(s, st);
    if (rc != DFU_OK) return rc;
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
    (void) cudaGetLastError();
    cudaStream_t st = as_stream(stream);
    // DFU_SOLVER_PATH=multi forces the one-kernel-per-phase path (used by the tests to cover both)
    const char* force = getenv("DFU_SOLVER_PATH");
    const bool multi = s->allreduce != nullptr || s->coop_blocks == 0 || (force && force[0] == 'm');
    if (!s->lists_sorted && (multi || !pattern_eligible(s)) && s->P > 0) {
        // the path changed after init_problem (hook set, environment): the float-order-dependent kernels want sorted lists
        const long ne = (long) s->P * 8;
        DFU_CUDA_OK(cudaMemsetAsync(s->tmp + s->N, 0, (size_t) s->N * sizeof(int), st));
        k_fill<<<div_up(ne, TPB), TPB, 0, st>>>(s->nbr, ne, s->tptr, s->tmp + s->N, s->tent, 0);
        DFU_LAUNCH_OK();
        k_sort_emit<<<min(div_up((long) s->N * 32, TPB), 65535), TPB, 0, st>>>(s->tptr, s->N, s->tent, s->wts, s->tv, s->tw);
        DFU_LAUNCH_OK();
        s->lists_sorted = true;
        s->pattern_ready = false;  // its per-entry slots referred to the old order
    }
    int rc = multi ? solve_multi_kernel(s, st) : solve_persistent(s, st);
    if (rc != DFU_OK) return rc;
    // write back ONCE: dg_se3 := DQ(0,0,0,t) * dg_se3 (opt_solver.cpp:270-285, node.cpp:19-23)
    rc = dfu_warpfield_update_translations(s->wf, s->vec, stream);
    if (prev != s->wf->device) cudaSetDevice(prev);
    return rc;
}

// CombinedSolver::updateHuberWeights (opt_solver.cpp:241-268): for node i the loop over its 8 neighbours overwrites
// h[i] every time, so the value that survives is the one of the LAST (8th nearest) neighbour j:
//   e = | T_i(dg_v[j]) - T_j(dg_v[j]) |,  h = e <= psi_reg ? 1 : psi_reg / e
// The reference computes it and never reads it (energy.t:76-77); provided for API completeness.
__global__ void k_huber(const float4* __restrict__ pos_w, const float4* __restrict__ real, const float4* __restrict__ dual,
                        const int32_t* __restrict__ nnbr, int N, float psi_reg, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int j = nnbr[(size_t) i * 8 + 7];
    const float4 pj = pos_w[j];
    const V3 c{pj.x, pj.y, pj.z};
    const V3 a = dq_transform_vertex(DQ{make_quat(real[i]), make_quat(dual[i])}, c);
    const V3 b = dq_transform_vertex(DQ{make_quat(real[j]), make_quat(dual[j])}, c);
    const float ex = fsub(a.x, b.x), ey = fsub(a.y, b.y), ez = fsub(a.z, b.z);
    const float e = __fsqrt_rn(fadd(fadd(fmul(ex, ex), fmul(ey, ey)), fmul(ez, ez)));
    out[i] = e <= psi_reg ? 1.f : __fdiv_rn(psi_reg, e);
}

int dfu_solver_huber_weights(const dfu_solver* s, float* huber, dfu_stream stream) {
    DFU_REQUIRE(s && huber, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(s->problem_ready, DFU_ERR_NOT_INIT, "no problem instance");
    k_huber<<<div_up(s->N, TPB), TPB, 0, as_stream(stream)>>>(s->wf->pos_w, s->wf->real, s->wf->dual, s->nnbr, s->N, s->prm.psi_reg, huber);
    DFU_LAUNCH_OK();
    return DFU_OK;
}

int dfu_solver_tukey_weights(const dfu_solver* s, float* tukey, dfu_stream stream) {
    DFU_REQUIRE(s && tukey, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(s->problem_ready, DFU_ERR_NOT_INIT, "no problem instance");
    DFU_CUDA_OK(cudaMemcpyAsync(tukey, s->theta, (size_t) s->P * sizeof(float), cudaMemcpyDeviceToDevice, as_stream(stream)));
    return DFU_OK;
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
    stats_host[3] = (double) (s->gn_steps_host >= 0 ? s->gn_steps_host : s->sc_host->gn_steps);
    DFU_REQUIRE(!(s->last_kernel == 3 && s->sc_host->spin_fail), DFU_ERR_CUDA,
                "persistent solver: an exchanged word never arrived (grid not co-resident?); result invalid");
    return DFU_OK;
}

}  // extern "C"
