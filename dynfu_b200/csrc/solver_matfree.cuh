// Solver: persistent cooperative kernels, versions 1 and 2 (matrix-free textbook PCG)
// (textually included by solver.cu inside its anonymous namespace -- one translation unit)
#pragma once

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

