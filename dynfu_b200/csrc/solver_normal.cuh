// Solver: persistent cooperative kernels, version 3 (explicit normal matrix + pipelined PCG) and its sparsity pattern
// (textually included by solver.cu inside its anonymous namespace -- one translation unit)
#pragma once

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

// Accumulators of one warp of version 3r: ACC_W low words + ACC_W high words (generic rows), or -- rows of at most 64 columns,
// the usual case -- four bank-rotated copies of 64 low + 64 high words and the three 64-bit sums of b.
constexpr int P3_XS_MAX = 1024;  // columns of the exchanged vector a CTA stages in shared memory (more: straight from L2)
constexpr int P3_FU = 2;  // entries per lane in flight in the fused sweep
constexpr int P3_ACC_WORDS = 640, P3_ACC_COPY = 72, P3_ACC_HI = 320, P3_ACC_B = 616;

// One sweep of node n's list (entries lo..hi of the transposed data graph) that accumulates BOTH b_n = sum tw * theta e (three
// 64-bit fixed-point sums, left at acc + P3_ACC_B) and the data part of row n of A (2^40 fixed point as 20 + 20 bit halves in
// 32-bit shared-memory atomics; slot s of the row is the sum of acc[s + 72 c] / acc[320 + s + 72 c] over the copies c).
// Measured (profiles/r02_solver_experiments.md): the loop is bound by shared-memory atomic wavefronts -- the 32 points of a
// batch share a handful of popular neighbours, 5.3 lanes per word per instruction on a single copy -- and by its two dependent
// L2 round trips.  So: FOUR copies of the accumulators, 72 words apart so that one slot lies in four different banks (lane & 3
// picks the copy; integer sums, any split gives the same bits); theta read from s4.w, the same 32-byte sector as theta e; and
// the point ids of the NEXT 128 entries fetched a batch ahead.
template <int FU>
DFU_DEV void assemble_row_fused(const int32_t* __restrict__ tv, const float* __restrict__ tw,
                                                const uint4* __restrict__ tslot, const float4* s4, const float* __restrict__ wts,
                                                int lo, int hi, int lane, unsigned* acc) {
    for (int j = lane; j < P3_ACC_WORDS; j += 32) acc[j] = 0u;
    unsigned* my_lo = acc + P3_ACC_COPY * (lane & 3);
    unsigned* my_hi = my_lo + P3_ACC_HI;
    int vn[FU];
    long long bx = 0, by = 0, bz = 0;
#pragma unroll
    for (int u = 0; u < FU; ++u) vn[u] = lo + lane + 32 * u < hi ? __ldcs(tv + lo + lane + 32 * u) : -1;
    __syncwarp();
    for (int e0 = lo + lane; e0 < hi; e0 += 32 * FU) {  // FU entries per lane in flight
        int v[FU];
        float c[FU];
        uint4 sl[FU];
        float4 se[FU];
        float wq[FU][8];
#pragma unroll
        for (int u = 0; u < FU; ++u) {
            v[u] = vn[u];
            const int vv = v[u] >= 0 ? v[u] : 0;
            const int ee = v[u] >= 0 ? e0 + 32 * u : lo;
            c[u] = __ldcs(tw + ee);  // (read once: streaming, the L1 is for the points' data shared by neighbouring rows)
            sl[u] = __ldcs(tslot + ee);
            se[u] = s4[vv];
            ld256(wts + 8 * (size_t) vv, wq[u]);
        }
#pragma unroll
        for (int u = 0; u < FU; ++u) vn[u] = e0 + 32 * FU + 32 * u < hi ? __ldcs(tv + e0 + 32 * FU + 32 * u) : -1;
#pragma unroll
        for (int u = 0; u < FU; ++u) {
            if (v[u] < 0) continue;
            bx += __float2ll_rn(c[u] * se[u].x * FIX_SCALE);
            by += __float2ll_rn(c[u] * se[u].y * FIX_SCALE);
            bz += __float2ll_rn(c[u] * se[u].z * FIX_SCALE);
            const float cc = c[u] * se[u].w;
            if (cc == 0.f) continue;
            const unsigned sw[4] = {sl[u].x, sl[u].y, sl[u].z, sl[u].w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const unsigned long long f = (unsigned long long) __float2ll_rn(cc * wq[u][k] * FIX_SCALE);
                const unsigned sidx = (k & 1) ? sw[k >> 1] >> 16 : sw[k >> 1] & 0xffffu;
                atomicAdd(&my_lo[sidx], (unsigned) f & 0xfffffu);
                atomicAdd(&my_hi[sidx], (unsigned) (f >> 20));
            }
        }
    }
    warp_sum_ll(bx); warp_sum_ll(by); warp_sum_ll(bz);
    if (lane == 0) {
        long long* out = reinterpret_cast<long long*>(acc + P3_ACC_B);
        out[0] = bx; out[1] = by; out[2] = bz;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(PTPB, 1) k_solve_persistent3(Problem pb, Pattern pt, SolveCtl ctl, Scalars* sc, unsigned* bar) {
    __shared__ double sh4[4 * (PTPB / 32)];
    __shared__ __align__(16) unsigned long long acc_sm[(PTPB / 32) * (P3_ACC_WORDS / 2)];  // per warp: ACC_W 64-bit words, or the layout of assemble_row_fused
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, gw = tid >> 5, nw = nthreads >> 5;
    const int nb = gridDim.x, N = pb.N;
    // every warp owns a contiguous block of rows (at most 32: N <= 65535 and >= 2368 resident warps): warp-per-row for the
    // row products, lane-per-row (coalesced float4 accesses) for the vector updates
    const int R = (N + nw - 1) / nw;
    const int row0 = min(N, gw * R), row1 = min(N, row0 + R);
    unsigned bar_target = 0;
    unsigned long long* acc = acc_sm + (threadIdx.x >> 5) * (P3_ACC_WORDS / 2);
    float4* S_r = pt.st;  // b of every row, handed from the warp that prepared the row to the lane that owns it in the PCG loop
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

    // lane l owns row row0 + l in the vector updates: its offset / length (fixed for the launch) stay in registers
    const int my_off = row0 + lane < row1 ? pt.rowptr[row0 + lane] : 0;
    const int my_len = row0 + lane < row1 ? pt.rowlen[row0 + lane] : 0;
    // y = A x for all rows of this warp, RQ rows in flight (their index / value loads are issued together, then their gathers:
    // two dependent L2 round trips per batch -- the row bounds come from the owning lanes by shuffle); the lane that owns row n
    // (lane == n - row0) receives the result
    constexpr int RQ = 8;
    auto spmv_block = [&](const float4* __restrict__ x, float& rx, float& ry, float& rz_) {
        rx = ry = rz_ = 0.f;
        for (int n0 = row0; n0 < row1; n0 += RQ) {
            int off[RQ], len[RQ], c[RQ][2];
            float v[RQ][2];
#pragma unroll
            for (int q = 0; q < RQ; ++q) {
                const int src = min(n0 + q - row0, 31);
                off[q] = __shfl_sync(0xffffffffu, my_off, src);
                len[q] = n0 + q < row1 ? __shfl_sync(0xffffffffu, my_len, src) : 0;
            }
#pragma unroll
            for (int q = 0; q < RQ; ++q)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int j = lane + 32 * u;
                    c[q][u] = j < len[q] ? pt.col[off[q] + j] : -1;
                    v[q][u] = j < len[q] ? pt.vals[off[q] + j] : 0.f;
                }
            float ax[RQ], ay[RQ], az[RQ];
#pragma unroll
            for (int q = 0; q < RQ; ++q) {
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
            for (int q = 0; q < RQ; ++q) {
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
                    const int off = pt.rowptr[n], len = pt.rowlen[n];
                    const int lo = pb.tptr[n], hi = pb.tptr[n + 1];
                    // the usual row (at most 64 columns, at most FIX_MAX_DEG points) when theta changed: b and the data part
                    // of the row in ONE sweep of the node's list (assemble_row_fused: four bank-rotated accumulator copies,
                    // native 32-bit shared-memory atomics, theta from s4.w); same integer sums as the general passes below
                    const bool fused = gn == 0 && hi - lo <= FIX_MAX_DEG && len <= 64;
                    if (pb.wreg2 > 0.f) {
                        node_gather_reg(pb, n, lane, pb.t, gx, gy, gz, cnt, e2);
                        gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
                        e2 = warp_sum(e2);
                    }
                    if (!fused) node_gather_data_fixed(pb, n, lane, ax, ay, az);  // already summed over the warp
                    PROF(3);
                    if (fused) {
                        unsigned* a32 = reinterpret_cast<unsigned*>(acc);
                        assemble_row_fused<P3_FU>(pb.tv, pb.tw, pt.tslot, pb.s4, pb.wts, lo, hi, lane, a32);
                        const long long* bsum = reinterpret_cast<const long long*>(a32 + P3_ACC_B);
                        ax = (float) ((double) bsum[0] * FIX_INV); ay = (float) ((double) bsum[1] * FIX_INV); az = (float) ((double) bsum[2] * FIX_INV);
                        for (int j = lane; j < len; j += 32) {
                            const unsigned* q = a32 + j;
                            const unsigned long long sl4 = (unsigned long long) q[0] + q[P3_ACC_COPY] + q[2 * P3_ACC_COPY] + q[3 * P3_ACC_COPY];
                            const unsigned long long sh4s = (unsigned long long) q[P3_ACC_HI] + q[P3_ACC_HI + P3_ACC_COPY] +
                                                            q[P3_ACC_HI + 2 * P3_ACC_COPY] + q[P3_ACC_HI + 3 * P3_ACC_COPY];
                            pt.vals[off + j] = pt.areg[off + j] + (float) ((double) ((sh4s << 20) + sl4) * FIX_INV);
                        }
                        __syncwarp();
                    } else if (gn == 0) {  // theta changed: data part of the row, ACC_W columns per pass
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
                    if (pb.wreg2 > 0.f) {
                        ax -= pb.wreg2 * gx; ay -= pb.wreg2 * gy; az -= pb.wreg2 * gz;
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
                // (the PCG vectors of the lane's row live in registers for the whole loop: r was written by whichever warp
                //  prepared the row, x is added to t at the end)
                float4 r_ = zero4, w_ = zero4, z_ = zero4, s_ = zero4, p_ = zero4, x_ = zero4;
                {
                    float wx, wy, wz;
                    spmv_block(pt.exch, wx, wy, wz);
                    const int n = row0 + lane;
                    if (n < row1) {
                        r_ = S_r[n];
                        w_ = make_float4(wx, wy, wz, 0.f);
                    }
                }
                // (one FP64 division between the totals and the update instead of three: 1 / gamma of the previous iteration is
                //  formed right after that iteration's update, and gamma / alpha_prev = beta * denom_prev; the row's diagonal and
                //  its reciprocals stay in registers across the iterations -- one row per lane)
                double gamma_prev = 0.0, inv_gamma_prev = 0.0, denom_prev = 0.0;
                float inv = 0.f;
                double invd = 0.0;
                if (row0 + lane < row1) {
                    const float D = pb.nbuf[3 * (size_t) N + row0 + lane];
                    inv = D > 0.f ? 1.f / D : 0.f;
                    invd = D > 0.f ? 1.0 / (double) D : 0.0;
                }
                PROF(8);
                for (int it = 0; it < ctl.linear_iter; ++it) {
                    const int buf = (it + 1) & 1;
                    float4* ex = pt.exch + (size_t) buf * N;
                    // (r,u), (w,u) with u = M^-1 r; m = M^-1 w goes to the exchange buffer
                    double g = 0.0, d = 0.0;
                    {
                        const int n = row0 + lane;  // one row per lane
                        if (n < row1) {
                            const float4 r = r_, w = w_;
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
                    const double beta = it > 0 ? gamma * inv_gamma_prev : 0.0;
                    const double denom = it > 0 ? delta - beta * beta * denom_prev : delta;
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
                            float4 r = r_, w = w_, z = z_, sv = s_, p = p_, x = x_;
                            z.x = __fmaf_rn(bf, z.x, nx); z.y = __fmaf_rn(bf, z.y, ny); z.z = __fmaf_rn(bf, z.z, nz);
                            sv.x = __fmaf_rn(bf, sv.x, w.x); sv.y = __fmaf_rn(bf, sv.y, w.y); sv.z = __fmaf_rn(bf, sv.z, w.z);
                            p.x = __fmaf_rn(bf, p.x, r.x * inv); p.y = __fmaf_rn(bf, p.y, r.y * inv); p.z = __fmaf_rn(bf, p.z, r.z * inv);
                            x.x = __fmaf_rn(af, p.x, x.x); x.y = __fmaf_rn(af, p.y, x.y); x.z = __fmaf_rn(af, p.z, x.z);
                            r.x = __fmaf_rn(-af, sv.x, r.x); r.y = __fmaf_rn(-af, sv.y, r.y); r.z = __fmaf_rn(-af, sv.z, r.z);
                            w.x = __fmaf_rn(-af, z.x, w.x); w.y = __fmaf_rn(-af, z.y, w.y); w.z = __fmaf_rn(-af, z.z, w.z);
                            r_ = r; w_ = w; z_ = z; s_ = sv; p_ = p; x_ = x;
                        }
                    }
                    PROF(12);
                    ++pcg_total;
                    gamma_prev = gamma;
                    inv_gamma_prev = 1.0 / gamma;
                    denom_prev = denom;
                }
                // t += x (row-local), then everybody needs the new t
                {
                    const int n = row0 + lane;
                    if (n < row1) {
                        pb.t[3 * (size_t) n] += x_.x; pb.t[3 * (size_t) n + 1] += x_.y; pb.t[3 * (size_t) n + 2] += x_.z;
                    }
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
    constexpr int ACC_WORDS = P3_ACC_WORDS, ACC_COPY = P3_ACC_COPY, ACC_HI = P3_ACC_HI;
    __shared__ __align__(16) unsigned acc_sm[NWARP * ACC_WORDS];  // (layout: assemble_row_fused)
    // The exchanged vector, staged: a CTA's rows reference a few hundred distinct columns (rows of neighbouring nodes share
    // most of theirs), but every lane fetching its own column from L2 is 106 k sector requests per PCG iteration chip-wide
    // on the 512 lines of one 64 KB vector -- measured 2.1 k cycles per iteration (profiles/r02_solver_experiments.md).  The
    // CTA's column list is built once per launch; each iteration the CTA fetches those columns once, in ascending order,
    // into shared memory (aliasing the assembly accumulators, idle during PCG) and the rows gather from there.
    __shared__ unsigned short xlist[P3_XS_MAX];
    __shared__ int xs_n;
    float4* xs = reinterpret_cast<float4*>(acc_sm);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, gw = tid >> 5, nw = nthreads >> 5;
    const int nb = gridDim.x, N = pb.N;
    unsigned bar_target = 0;
    unsigned* acc_lo = acc_sm + wib * ACC_WORDS;
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
    float rv[P3_R][P3_LE];  // (the regularisation values of the same entries are re-read from pt.areg when a row is assembled)
    float rD[P3_R], rinv[P3_R];
    double rinvd[P3_R];
    float s_r[P3_R], s_w[P3_R], s_z[P3_R], s_s[P3_R], s_p[P3_R], s_x[P3_R], s_t[P3_R];  // coordinate `lane` (lanes 0..2)
#pragma unroll
    for (int r = 0; r < P3_R; ++r) {
        // (row r of a warp is gw + r * nw: a node early in the order paired with one nw further on.  Measured alternatives, both
        //  slower at N = 4096: second rows dealt round-robin over the CTAs for equal row counts per SM, 0.328 vs 0.302 ms -- the
        //  L1 reuse of the points shared by neighbouring rows is lost; contiguous chunks of N / gridDim.x rows per CTA, 0.335 ms
        //  -- list lengths vary smoothly along the node order, and the strided pairing is what balances them)
        const int n = gw + r * nw;
        rn[r] = n < N ? n : -1;
        roff[r] = rlen[r] = rds[r] = 0;
        rD[r] = rinv[r] = 0.f;
        rinvd[r] = 0.0;
        s_r[r] = s_w[r] = s_z[r] = s_s[r] = s_p[r] = s_x[r] = s_t[r] = 0.f;
#pragma unroll
        for (int u = 0; u < P3_LE; ++u) {
            rc[r][u] = -1;
            rv[r][u] = 0.f;
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
                }
            }
            if (lane < 3) pb.t[3 * (size_t) n + lane] = 0.f;  // unknowns := 0 (opt_solver.cpp:192-193)
            if (lane == 0) pt.t4[n] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    if (pt.wf_flags && tid == 0) {  // (recomputed with the write-back at the end)
        pt.wf_flags[0] = 1; pt.wf_flags[1] = 0; pt.wf_flags[2] = 0;
    }
    bool staged;
    {
        // bitmap of the columns the CTA's register entries reference -> ascending list, registers hold the position in it
        unsigned* bm = acc_sm;
        unsigned* bpre = acc_sm + 256;
        const int nwb = (N + 31) >> 5;  // <= 148: N <= P3_R * 16 * gridDim.x
        for (int w = threadIdx.x; w < nwb; w += blockDim.x) bm[w] = 0u;
        __syncthreads();
#pragma unroll
        for (int r = 0; r < P3_R; ++r)
#pragma unroll
            for (int u = 0; u < P3_LE; ++u)
                if (rc[r][u] >= 0) atomicOr(&bm[rc[r][u] >> 5], 1u << (rc[r][u] & 31));
        __syncthreads();
        if (wib == 0) {
            int base = 0;
            for (int w0 = 0; w0 < nwb; w0 += 32) {
                const int w = w0 + lane;
                const int c = w < nwb ? __popc(bm[w]) : 0;
                int inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                if (w < nwb) bpre[w] = (unsigned) (base + inc - c);
                base += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) xs_n = base;
        }
        __syncthreads();
        staged = xs_n <= P3_XS_MAX && nwb <= 256;
        if (staged) {
            for (int w = threadIdx.x; w < nwb; w += blockDim.x) {
                unsigned bits = bm[w];
                int j = (int) bpre[w];
                while (bits) {
                    xlist[j++] = (unsigned short) (32 * w + __ffs(bits) - 1);
                    bits &= bits - 1;
                }
            }
#pragma unroll
            for (int r = 0; r < P3_R; ++r)
#pragma unroll
                for (int u = 0; u < P3_LE; ++u)
                    if (rc[r][u] >= 0)
                        rc[r][u] = (int) bpre[rc[r][u] >> 5] + __popc(bm[rc[r][u] >> 5] & ((1u << (rc[r][u] & 31)) - 1u));
        }
        __syncthreads();
    }
    const int xs_count = staged ? xs_n : 0;
    // the CTA's columns of x -> shared memory (every thread of the CTA calls this; ends with a CTA barrier)
    auto stage_x = [&](const float4* x) {
        if (staged) {
            for (int i = threadIdx.x; i < xs_count; i += PTPB) xs[i] = __ldcg(x + xlist[i]);
            __syncthreads();
        }
    };
    GRID_SYNC();

    // row product with the exchanged vector x (float4 per row): registers first, entries beyond 32 * P3_LE from L2
    auto spmv = [&](int r, const float4* x) -> float {
        float ax = 0.f, ay = 0.f, az = 0.f;
        float4 m[P3_LE];
#pragma unroll
        for (int u = 0; u < P3_LE; ++u)
            m[u] = rc[r][u] >= 0 ? (staged ? xs[rc[r][u]] : __ldcg(x + rc[r][u])) : make_float4(0.f, 0.f, 0.f, 0.f);
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
            const double e2_local = phase_point_residual_t4(pb, pt.t4, gn == 0, tid, nthreads);
            PROF(1);
            GRID_SYNC();
            PROF(2);
            double rz = 0.0, er = 0.0;
#pragma unroll
            for (int r = 0; r < P3_R; ++r) {
                if (rn[r] < 0) continue;  // uniform over the warp
                const int n = rn[r];
                float ax, ay, az, gx = 0.f, gy = 0.f, gz = 0.f, cnt = 0.f, e2 = 0.f;
                // b_n = sum tw * theta e in 2^40 fixed point (independent of the order of the node's list) and, when theta
                // changed, the data part of row n of A (fixed point in shared memory), ACC_W columns per pass.  The usual row
                // does both in ONE sweep of the node's list: theta travels in s4.w, the same 32-byte sector as theta e.
                const bool assemble = gn == 0;
                const int lo = pb.tptr[n], hi = pb.tptr[n + 1];
                const bool wide = hi - lo > FIX_MAX_DEG;  // too many contributions for the split words: 64-bit atomics
                const bool fused = assemble && !wide && rlen[r] <= 32 * P3_LE;
                if (!fused) node_gather_data_fixed(pb, n, lane, ax, ay, az);
                PROF(3);
                if (assemble) {
                    unsigned long long* acc64 = reinterpret_cast<unsigned long long*>(acc_lo);
                    if (fused) {
                        // the usual row (all of it in registers): one sweep, out of line so that its 80 registers of loads in flight
                        // are not live across the PCG loop
                        float areg_u[P3_LE];  // (fetched now: a dependent L2 round trip after the sweep otherwise)
#pragma unroll
                        for (int u = 0; u < P3_LE; ++u) areg_u[u] = rc[r][u] >= 0 ? pt.areg[roff[r] + lane + 32 * u] : 0.f;
                        assemble_row_fused<P3_FU>(pb.tv, pb.tw, pt.tslot, pb.s4, pb.wts, lo, hi, lane, acc_lo);
                        const long long bx = reinterpret_cast<const long long*>(acc_lo + P3_ACC_B)[0],
                                        by = reinterpret_cast<const long long*>(acc_lo + P3_ACC_B)[1],
                                        bz = reinterpret_cast<const long long*>(acc_lo + P3_ACC_B)[2];
                        ax = (float) ((double) bx * FIX_INV); ay = (float) ((double) by * FIX_INV); az = (float) ((double) bz * FIX_INV);
                        __syncwarp();
#pragma unroll
                        for (int u = 0; u < P3_LE; ++u)
                            if (rc[r][u] >= 0) {
                                const unsigned* a = acc_lo + lane + 32 * u;
                                const unsigned long long sl4 = (unsigned long long) a[0] + a[ACC_COPY] + a[2 * ACC_COPY] + a[3 * ACC_COPY];
                                const unsigned long long sh4 = (unsigned long long) a[ACC_HI] + a[ACC_HI + ACC_COPY] +
                                                               a[ACC_HI + 2 * ACC_COPY] + a[ACC_HI + 3 * ACC_COPY];
                                rv[r][u] = areg_u[u] + (float) ((double) ((sh4 << 20) + sl4) * FIX_INV);
                            }
                        __syncwarp();
                    } else
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
                                    rv[r][u] = pt.areg[roff[r] + lane + 32 * u] + (wide ? (float) ((double) acc64[lane + 32 * u] * FIX_INV)
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
                    node_gather_reg_t4(pb, n, lane, pt.t4, gx, gy, gz, e2);
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
                stage_x(pt.exch);
#pragma unroll
                for (int r = 0; r < P3_R; ++r) {
                    if (rn[r] < 0) continue;
                    const float w = spmv(r, pt.exch);  // w0 = A u0
                    s_w[r] = lane < 3 ? w : 0.f;
                    s_z[r] = s_s[r] = s_p[r] = 0.f;
                }
                // (1 / gamma of the previous iteration is formed right after that iteration's update, off the critical path, and
                //  gamma / alpha_prev = beta * denom_prev: one FP64 division between the totals and the update instead of three)
                double gamma_prev = 0.0, inv_gamma_prev = 0.0, denom_prev = 0.0;
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
                    // only lanes 0..2 hold terms: two xor steps bring their sum to lane 0
                    g += __shfl_xor_sync(0xffffffffu, g, 1); d += __shfl_xor_sync(0xffffffffu, d, 1);
                    g += __shfl_xor_sync(0xffffffffu, g, 2); d += __shfl_xor_sync(0xffffffffu, d, 2);
                    if (lane == 0) {  // (shw was last read before the previous grid barrier)
                        shw[wib] = g; shw[NWARP + wib] = d;
                    }
                    __syncthreads();  // the CTA's m and partial terms are written
                    PROF(9);
                    if (wib == 0) {
                        // CTA total in a fixed order (xor tree over the 16 warps), published, then the grid barrier proper:
                        // the release covers every store of the CTA made before the __syncthreads above
                        double tg = lane < NWARP ? shw[lane] : 0.0, td = lane < NWARP ? shw[NWARP + lane] : 0.0;
#pragma unroll
                        for (int o = NWARP / 2; o > 0; o >>= 1) {
                            tg += __shfl_xor_sync(0xffffffffu, tg, o);
                            td += __shfl_xor_sync(0xffffffffu, td, o);
                        }
                        if (lane == 0) {
                            partf[blockIdx.x] = make_float2((float) tg, (float) td);
                            bar_target += (unsigned) nb;
                            unsigned seen;
                            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
                            do {
                                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
                            } while (seen < bar_target);
                        }
                    }
                    __syncthreads();
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
                    if (!last) stage_x(ex);
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
                    const double beta = it > 0 ? gamma * inv_gamma_prev : 0.0;
                    const double denom = it > 0 ? delta - beta * beta * denom_prev : delta;
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
                    inv_gamma_prev = 1.0 / gamma;
                    denom_prev = denom;
                }
            }
#pragma unroll
            for (int r = 0; r < P3_R; ++r)
                if (rn[r] >= 0 && lane < 3) {
                    s_t[r] += s_x[r];
                    pb.t[3 * (size_t) rn[r] + lane] = s_t[r];
                    reinterpret_cast<float*>(pt.t4 + rn[r])[lane] = s_t[r];
                }
            ++gn_total;
            PROF(13);
            GRID_SYNC();
            PROF(14);
        }
    }
    // ---- final energy at the solution, Tukey weights of the last outer iteration -------------------------
    {
        const double e2 = phase_point_residual_t4(pb, pt.t4, first, tid, nthreads);  // no GN step ran: weights at t = 0
        double er = 0.0;
        if (pb.wreg2 > 0.f) {
#pragma unroll
            for (int r = 0; r < P3_R; ++r) {
                if (rn[r] < 0) continue;
                float gx, gy, gz, r2;
                node_gather_reg_t4(pb, rn[r], lane, pt.t4, gx, gy, gz, r2);
                r2 = warp_sum(r2);
                if (lane == 0) er += (double) pb.wreg2 * r2;
            }
        }
        // write back ONCE: dg_se3 := DQ(0,0,0,t) * dg_se3 (opt_solver.cpp:270-285, node.cpp:19-23), and the warp field's
        // flags (see node_flags_kernel) -- every t is final and visible since the last grid barrier
        if (pt.wf_real && tid < ((N + 31) & ~31)) {
            int ok = 1;
            float mw = 0.f, md = 0.f;
            if (tid < N) {
                const DQ inc = dq_from_translation(pb.t[3 * (size_t) tid], pb.t[3 * (size_t) tid + 1], pb.t[3 * (size_t) tid + 2]);
                const DQ old{make_quat(pt.wf_real[tid]), make_quat(pt.wf_dual[tid])};
                const DQ res = dq_mul(inc, old);
                const float4 r = to_float4(res.real), d = to_float4(res.dual);
                pt.wf_real[tid] = r;
                pt.wf_dual[tid] = d;
                ok = (r.x == 1.f && r.y == 0.f && r.z == 0.f && r.w == 0.f && d.x == 0.f);
                mw = pt.wf_pos_w[tid].w;
                md = fmaxf(fabsf(d.y), fmaxf(fabsf(d.z), fabsf(d.w)));
            }
            ok = __all_sync(0xffffffffu, ok);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
                md = fmaxf(md, __shfl_xor_sync(0xffffffffu, md, o));
            }
            if (lane == 0) {
                if (!ok) atomicAnd(&pt.wf_flags[0], 0);
                atomicMax(reinterpret_cast<unsigned*>(&pt.wf_flags[1]), __float_as_uint(mw));  // non-negative floats
                atomicMax(reinterpret_cast<unsigned*>(&pt.wf_flags[2]), __float_as_uint(md));
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
    // the last CTA out clears the barrier words for the next launch (everybody is past the final barrier when it leaves)
    if (threadIdx.x == 0) {
        const unsigned out = atomicAdd(bar + 1, 1u);
        if (out == (unsigned) nb - 1u) {
            bar[0] = 0u;
            bar[1] = 0u;
        }
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
        // (both sweeps of the node's list keep four entries per lane in flight: the kernel is one wave of warps, each a chain of
        //  dependent point-id -> neighbour-list loads; the second sweep finds the lists in L1)
        for (int e0 = lo + lane; e0 < hi; e0 += 128) {
            int v[4];
            int nq[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = e0 + 32 * u < hi ? pb.tv[e0 + 32 * u] : -1;
#pragma unroll
            for (int u = 0; u < 4; ++u) ld256(pb.nbr + 8 * (size_t) (v[u] >= 0 ? v[u] : 0), nq[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (v[u] < 0) continue;
                const int (&nbk)[8] = nq[u];
#pragma unroll
                for (int k = 0; k < 8; ++k) atomicOr(&bm[nbk[k] >> 5], 1u << (nbk[k] & 31));
            }
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
        for (int e0 = lo + lane; e0 < hi; e0 += 128) {
            int v[4];
            int nq[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = e0 + 32 * u < hi ? pb.tv[e0 + 32 * u] : -1;
#pragma unroll
            for (int u = 0; u < 4; ++u) ld256(pb.nbr + 8 * (size_t) (v[u] >= 0 ? v[u] : 0), nq[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (v[u] < 0) continue;
                const int (&nbk)[8] = nq[u];
                unsigned sk[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) sk[k] = slot_of(nbk[k]);
                tslot[e0 + 32 * u] = make_uint4(sk[0] | (sk[1] << 16), sk[2] | (sk[3] << 16), sk[4] | (sk[5] << 16), sk[6] | (sk[7] << 16));
            }
        }
        __syncwarp();
    }
}

