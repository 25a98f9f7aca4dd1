// Solver: problem description, reductions, grid barrier, the phase functions shared by every execution path
// (textually included by solver.cu inside its anonymous namespace -- one translation unit)
#pragma once


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
    // version 3r writes the solution back itself (dg_se3 := DQ(0,0,0,t) * dg_se3 and the warp field's flags) when these are set
    float4* wf_real;
    float4* wf_dual;
    const float4* wf_pos_w;
    int* wf_flags;
    float4* t4;  // version 3r: the unknowns once more as one float4 per node (a point's 8 gathers are 8 loads instead of 24)
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
    ld256(nbr + 8 * (size_t) v, nb);  // (the graph of a frame is immutable while the solver kernels run)
    ld256(wts + 8 * (size_t) v, w);
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

// the same with the unknowns as one float4 per node: the residual pass is bound by the L1 requests of its gathers (24 scalar loads
// per point from the float[3N] array), a float4 per neighbour is 8
DFU_DEV double phase_point_residual_t4(const Problem& pb, const float4* __restrict__ t4, bool update_tukey, int tid, int nthreads) {
    double e2 = 0.0;
    for (int v = tid; v < pb.P; v += nthreads) {
        int nb[8];
        float w[8];
        load8(pb.nbr, pb.wts, v, nb, w);
        float4 xk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) xk[k] = t4[nb[k]];
        float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            sx = __fmaf_rn(w[k], xk[k].x, sx);
            sy = __fmaf_rn(w[k], xk[k].y, sy);
            sz = __fmaf_rn(w[k], xk[k].z, sz);
        }
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

// node_gather_reg on the float4 copy of the unknowns (one load per neighbour instead of three; same terms, same order)
DFU_DEV void node_gather_reg_t4(const Problem& pb, int n, int lane, const float4* __restrict__ x4, float& gx, float& gy, float& gz,
                                float& e2) {
    gx = gy = gz = e2 = 0.f;
    const float4 xn = x4[n];
    const int lo = pb.rin_ptr[n], hi = pb.rin_ptr[n + 1];
    for (int j = lane; j < 8 + (hi - lo); j += 32) {
        const bool out = j < 8;
        const int m = out ? pb.nnbr[(size_t) n * 8 + j] : pb.rin[lo + j - 8];
        if (m == n) continue;
        const float4 xm = x4[m];
        const float d0 = xn.x - xm.x, d1 = xn.y - xm.y, d2 = xn.z - xm.z;
        gx += d0; gy += d1; gz += d2;
        if (out) e2 += d0 * d0 + d1 * d1 + d2 * d2;
    }
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

