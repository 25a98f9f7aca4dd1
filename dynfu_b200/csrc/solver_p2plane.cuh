// Solver, north-star extension P2PLANE_SE3 (BASELINE.json north_star (4)): point-to-plane data term with one rigid increment
// per node.  No reference implementation exists (energy.t is translation-only, point-to-point):
// the yardstick is the double-precision CPU restatement used by the tests (orc_solve_p2plane), itself pinned against
// scipy.optimize.least_squares.
//
//   X_k in SE(3) per node (identity at the start), w^_vk = w_vk / sum_j w_vj
//   p_v(X) = sum_k w^_vk X_k c_v
//   E(X)   = sum_v theta_v (n_v . (p_v - l_v))^2 + w_reg^2 sum_n sum_{m in nbr(n), m != n} | X_n g_m - X_m g_m |^2
//   GN step X_k <- exp(xi_k) X_k with rows sqrt(theta) w^_vk [ (X_k c_v) x n_v ; n_v ], solved by block-Jacobi (6x6) PCG.
//   Per-node J^T J blocks and J^T r are GATHERS over the transposed graph (one warp per node, fixed xor-shuffle tree): no
//   atomics, reproducible; reductions in double.  One kernel per phase (this mode is not latency-tuned yet).
// (textually included by solver.cu inside its anonymous namespace -- one translation unit)
#pragma once

struct P2PProblem {
    int N, P;
    const int32_t* nbr;     // P*8
    const float* wts;       // P*8 un-normalised Gaussian weights
    const float *canon, *live, *nrm;
    const int* tptr;        // transposed graph: node -> points (sorted by point)
    const int32_t* tv;
    unsigned char* tk;      // per transposed entry: slot of the node in its point's neighbour list
    const int32_t* nnbr;    // N*8 regularisation out-edges
    const int* rin_ptr;     // in-edges (sources, ascending)
    const int32_t* rin;
    const float4* pos_w;
    float wreg2, tukey_offset, psi_data;
    // per point
    float* wn;     // P*8
    float* jac;    // P*8*6
    float* e;      // P
    float* theta;  // P
    float* sv;     // P: theta * (J x)
    // per node
    float* X;      // N*12 (R row-major, t)
    float* G;      // N*8*6: X_n g_m | X_m g_m for out-edge (n, i)
    float *b, *x, *r, *z, *p, *q;  // N*6
    float* L;      // N*36: Cholesky factor of the diagonal block (row-major lower), L[0] <= 0: singular block
    double* part;  // 4 * MAX_PARTIALS
};

DFU_DEV void p2p_apply(const float* X, float cx, float cy, float cz, float& ox, float& oy, float& oz) {
    ox = X[0] * cx + X[1] * cy + X[2] * cz + X[9];
    oy = X[3] * cx + X[4] * cy + X[5] * cz + X[10];
    oz = X[6] * cx + X[7] * cy + X[8] * cz + X[11];
}

__global__ void __launch_bounds__(TPB) kp_init(P2PProblem pb, Scalars* sc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        sc->rz[0] = sc->rz[1] = 0.0;
        sc->rz_ref = -1.0;
        sc->E = sc->E0 = 0.0;
        sc->done_it = INT_MAX;
        sc->pcg_iters = sc->gn_steps = 0;
        sc->first = 1;
        sc->spin_fail = 0;
    }
    if (i < pb.N) {
#pragma unroll
        for (int k = 0; k < 12; ++k) pb.X[12 * (size_t) i + k] = (k == 0 || k == 4 || k == 8) ? 1.f : 0.f;
    }
    if (i < pb.P) {
        float w[8], s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            w[k] = pb.wts[8 * (size_t) i + k];
            s += w[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) pb.wn[8 * (size_t) i + k] = s > 0.f ? w[k] / s : 0.f;
    }
}

// linearisation point: p_v, e_v = n.(p - l), the 8 Jacobian 6-vectors; optional Tukey update; partial sum of theta e^2
__global__ void __launch_bounds__(TPB) kp_linearise(P2PProblem pb, int update_tukey) {
    __shared__ double sh[TPB / 32];
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    double e2 = 0.0;
    if (v < pb.P) {
        const float cx = pb.canon[3 * (size_t) v], cy = pb.canon[3 * (size_t) v + 1], cz = pb.canon[3 * (size_t) v + 2];
        const float nx = pb.nrm[3 * (size_t) v], ny = pb.nrm[3 * (size_t) v + 1], nz = pb.nrm[3 * (size_t) v + 2];
        float px = 0.f, py = 0.f, pz = 0.f, sw = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float w = pb.wn[8 * (size_t) v + k];
            float qx, qy, qz;
            p2p_apply(pb.X + 12 * (size_t) pb.nbr[8 * (size_t) v + k], cx, cy, cz, qx, qy, qz);
            px += w * qx; py += w * qy; pz += w * qz;
            sw += w;
            float* a = pb.jac + ((size_t) v * 8 + k) * 6;
            a[0] = w * (qy * nz - qz * ny); a[1] = w * (qz * nx - qx * nz); a[2] = w * (qx * ny - qy * nx);
            a[3] = w * nx; a[4] = w * ny; a[5] = w * nz;
        }
        const float dx = px - pb.live[3 * (size_t) v], dy = py - pb.live[3 * (size_t) v + 1], dz = pz - pb.live[3 * (size_t) v + 2];
        const float e = nx * dx + ny * dy + nz * dz;
        pb.e[v] = e;
        float th;
        if (update_tukey) {
            th = sw > 0.f ? tukey_biweight(pb.tukey_offset, pb.psi_data, -dx, -dy, -dz) : 0.f;
            pb.theta[v] = th;
        } else {
            th = pb.theta[v];
        }
        e2 = (double) th * (double) e * (double) e;
    }
    const double bs = block_sum(e2, sh);
    if (threadIdx.x == 0) pb.part[blockIdx.x] = bs;
}

// X_n g_m and X_m g_m of every out-edge; partial regularisation energy
__global__ void __launch_bounds__(TPB) kp_edges(P2PProblem pb) {
    __shared__ double sh[TPB / 32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // edge id = n * 8 + k
    double er = 0.0;
    if (i < pb.N * 8) {
        const int n = i >> 3, m = pb.nnbr[i];
        const float4 g = pb.pos_w[m];
        float* G = pb.G + 6 * (size_t) i;
        p2p_apply(pb.X + 12 * (size_t) n, g.x, g.y, g.z, G[0], G[1], G[2]);
        p2p_apply(pb.X + 12 * (size_t) m, g.x, g.y, g.z, G[3], G[4], G[5]);
        if (m != n) {
            const double dx = (double) G[0] - G[3], dy = (double) G[1] - G[4], dz = (double) G[2] - G[5];
            er = (double) pb.wreg2 * (dx * dx + dy * dy + dz * dz);
        }
    }
    const double bs = block_sum(er, sh);
    if (threadIdx.x == 0) pb.part[MAX_PARTIALS + blockIdx.x] = bs;
}

// slot of node n in the neighbour list of point v
DFU_DEV int p2p_slot(const P2PProblem& pb, int v, int n) {
    int k = 0;
#pragma unroll
    for (int j = 1; j < 8; ++j) k = pb.nbr[8 * (size_t) v + j] == n ? j : k;
    return k;
}

__global__ void __launch_bounds__(TPB) kp_slots(P2PProblem pb) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int n = gw; n < pb.N; n += nw)
        for (int j = pb.tptr[n] + lane; j < pb.tptr[n + 1]; j += 32) pb.tk[j] = (unsigned char) p2p_slot(pb, pb.tv[j], n);
}

// regularisation part of (J^T J x)_n for node n (lane-parallel over its out- and in-edges), reduced over the warp
DFU_DEV void p2p_reg_apply(const P2PProblem& pb, int n, int lane, const float* x, float (&acc)[6]) {
    const int lo = pb.rin_ptr[n], hi = pb.rin_ptr[n + 1];
    for (int j = lane; j < 8 + 8 * (hi - lo); j += 32) {
        // out-edges: (n, j); in-edges: every out-edge of a source `src` that points at n
        const bool out = j < 8;
        const int src = out ? n : pb.rin[lo + (j - 8) / 8];
        const int i = out ? j : (j - 8) & 7;
        const int m = pb.nnbr[(size_t) src * 8 + i];
        if (m == src || (!out && (m != n || src == n))) continue;
        const float* G = pb.G + 6 * ((size_t) src * 8 + i);
        const float* xs = x + 6 * (size_t) src;
        const float* xm = x + 6 * (size_t) m;
        const float r0 = (xs[1] * G[2] - xs[2] * G[1]) + xs[3] - (xm[1] * G[5] - xm[2] * G[4]) - xm[3];
        const float r1 = (xs[2] * G[0] - xs[0] * G[2]) + xs[4] - (xm[2] * G[3] - xm[0] * G[5]) - xm[4];
        const float r2 = (xs[0] * G[1] - xs[1] * G[0]) + xs[5] - (xm[0] * G[4] - xm[1] * G[3]) - xm[5];
        const float* Gk = out ? G : G + 3;
        const float sg = out ? pb.wreg2 : -pb.wreg2;
        acc[0] += sg * (Gk[1] * r2 - Gk[2] * r1);
        acc[1] += sg * (Gk[2] * r0 - Gk[0] * r2);
        acc[2] += sg * (Gk[0] * r1 - Gk[1] * r0);
        acc[3] += sg * r0; acc[4] += sg * r1; acc[5] += sg * r2;
    }
}

// per node: b = -J^T r0, the 6x6 diagonal block and its Cholesky factor; PCG start x = 0, r = b, z = M^-1 b, p = z; partial r.z
__global__ void __launch_bounds__(TPB) kp_assemble(P2PProblem pb) {
    __shared__ double sh[TPB / 32];
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    double rz = 0.0;
    for (int n = gw; n < pb.N; n += nw) {
        double b[6] = {0, 0, 0, 0, 0, 0}, M[21];
#pragma unroll
        for (int i = 0; i < 21; ++i) M[i] = 0.0;
        for (int j = pb.tptr[n] + lane; j < pb.tptr[n + 1]; j += 32) {
            const int v = pb.tv[j];
            const float* a = pb.jac + ((size_t) v * 8 + pb.tk[j]) * 6;
            const double th = pb.theta[v], te = th * (double) pb.e[v];
            int idx = 0;
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                b[r] -= te * a[r];
#pragma unroll
                for (int c = 0; c <= r; ++c) M[idx++] += th * (double) a[r] * (double) a[c];
            }
        }
        // regularisation: rhs and diagonal block of the edges touching n
        if (pb.wreg2 > 0.f) {
            const int lo = pb.rin_ptr[n], hi = pb.rin_ptr[n + 1];
            for (int j = lane; j < 8 + 8 * (hi - lo); j += 32) {
                const bool out = j < 8;
                const int src = out ? n : pb.rin[lo + (j - 8) / 8];
                const int i = out ? j : (j - 8) & 7;
                const int m = pb.nnbr[(size_t) src * 8 + i];
                if (m == src || (!out && (m != n || src == n))) continue;
                const float* G = pb.G + 6 * ((size_t) src * 8 + i);
                const double r0 = (double) G[0] - G[3], r1 = (double) G[1] - G[4], r2 = (double) G[2] - G[5];
                const float* Gk = out ? G : G + 3;
                const double gx = Gk[0], gy = Gk[1], gz = Gk[2];
                const double sg = out ? (double) pb.wreg2 : -(double) pb.wreg2, w2 = pb.wreg2;
                b[0] -= sg * (gy * r2 - gz * r1); b[1] -= sg * (gz * r0 - gx * r2); b[2] -= sg * (gx * r1 - gy * r0);
                b[3] -= sg * r0; b[4] -= sg * r1; b[5] -= sg * r2;
                // J^T J = [ K^T K  K ; -K  I ],  K = [Gk]x; lower triangle, row-major packed (r, c <= r)
                M[0] += w2 * (gy * gy + gz * gz);
                M[1] += w2 * (-gx * gy); M[2] += w2 * (gx * gx + gz * gz);
                M[3] += w2 * (-gx * gz); M[4] += w2 * (-gy * gz); M[5] += w2 * (gx * gx + gy * gy);
                // rows 3..5 (tau) x cols 0..2 (omega): -K
                M[6] += 0.0;        M[7] += w2 * gz;   M[8] += w2 * (-gy);  M[9] += w2;
                M[10] += w2 * (-gz); M[11] += 0.0;      M[12] += w2 * gx;    M[14] += w2;
                M[15] += w2 * gy;   M[16] += w2 * (-gx); M[17] += 0.0;       M[20] += w2;
            }
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) b[r] = warp_sum(b[r]);
#pragma unroll
        for (int i = 0; i < 21; ++i) M[i] = warp_sum(M[i]);
        if (lane == 0) {
            // Cholesky of the 6x6 block (double), stored as float L (row-major lower); not positive definite -> L[0] = 0
            double Lm[21];
            bool ok = true;
            int idx = 0;
            for (int i = 0; i < 6 && ok; ++i)
                for (int j = 0; j <= i; ++j) {
                    double s = M[i * (i + 1) / 2 + j];
                    for (int k = 0; k < j; ++k) s -= Lm[i * (i + 1) / 2 + k] * Lm[j * (j + 1) / 2 + k];
                    if (i == j) {
                        if (!(s > 0.0)) {
                            ok = false;
                            break;
                        }
                        Lm[idx] = sqrt(s);
                    } else {
                        Lm[idx] = s / Lm[j * (j + 1) / 2 + j];
                    }
                    ++idx;
                }
            float* Lo = pb.L + 21 * (size_t) n;
            for (int i = 0; i < 21; ++i) Lo[i] = ok ? (float) Lm[i] : 0.f;
            double z[6] = {0, 0, 0, 0, 0, 0};
            if (ok) {
                double y[6];
                for (int i = 0; i < 6; ++i) {
                    double s = b[i];
                    for (int k = 0; k < i; ++k) s -= Lm[i * (i + 1) / 2 + k] * y[k];
                    y[i] = s / Lm[i * (i + 1) / 2 + i];
                }
                for (int i = 5; i >= 0; --i) {
                    double s = y[i];
                    for (int k = i + 1; k < 6; ++k) s -= Lm[k * (k + 1) / 2 + i] * z[k];
                    z[i] = s / Lm[i * (i + 1) / 2 + i];
                }
            }
            for (int r = 0; r < 6; ++r) {
                const size_t i = 6 * (size_t) n + r;
                pb.b[i] = (float) b[r]; pb.r[i] = (float) b[r]; pb.x[i] = 0.f;
                pb.z[i] = (float) z[r]; pb.p[i] = (float) z[r];
                rz += b[r] * z[r];
            }
        }
    }
    const double bs = block_sum(rz, sh);
    if (threadIdx.x == 0) pb.part[2 * MAX_PARTIALS + blockIdx.x] = bs;
}

// z = M^-1 r with the stored Cholesky factor
DFU_DEV void p2p_precond(const float* L, const float (&r)[6], float (&z)[6]) {
    if (!(L[0] > 0.f)) {
#pragma unroll
        for (int i = 0; i < 6; ++i) z[i] = 0.f;
        return;
    }
    float y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        float s = r[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s -= L[i * (i + 1) / 2 + k] * y[k];
        y[i] = s / L[i * (i + 1) / 2 + i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        float s = y[i];
#pragma unroll
        for (int k = i + 1; k < 6; ++k) s -= L[k * (k + 1) / 2 + i] * z[k];
        z[i] = s / L[i * (i + 1) / 2 + i];
    }
}

__global__ void __launch_bounds__(TPB) kp_init_scalars(P2PProblem pb, Scalars* sc, int nblk_p, int nblk_e, int nblk_w, double tol2) {
    __shared__ double sh[TPB / 32];
    if (blockIdx.x != 0) return;
    const double ed = sum_partials(pb.part, nblk_p, sh);
    const double er = sum_partials(pb.part + MAX_PARTIALS, nblk_e, sh);
    const double rz = sum_partials(pb.part + 2 * MAX_PARTIALS, nblk_w, sh);
    if (threadIdx.x == 0) {
        sc->E = ed + er;
        if (sc->first) {
            sc->E0 = sc->E;
            sc->first = 0;
        }
        if (sc->rz_ref < 0.0) sc->rz_ref = rz;
        sc->rz[0] = rz;
        sc->done_it = (!(rz > 0.0) || rz <= tol2 * sc->rz_ref) ? 0 : INT_MAX;
    }
}
// energy only (after the last update)
__global__ void __launch_bounds__(TPB) kp_final_energy(P2PProblem pb, Scalars* sc, int nblk_p, int nblk_e) {
    __shared__ double sh[TPB / 32];
    if (blockIdx.x != 0) return;
    const double ed = sum_partials(pb.part, nblk_p, sh);
    const double er = sum_partials(pb.part + MAX_PARTIALS, nblk_e, sh);
    if (threadIdx.x == 0) {
        sc->E = ed + er;
        if (sc->first) {
            sc->E0 = sc->E;
            sc->first = 0;
        }
    }
}

__global__ void __launch_bounds__(TPB) kp_point_apply(P2PProblem pb, const Scalars* sc, int it) {
    if (it >= sc->done_it) return;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= pb.P) return;
    const float th = pb.theta[v];
    float acc = 0.f;
    if (th != 0.f) {
        const float4* a4 = reinterpret_cast<const float4*>(pb.jac + (size_t) v * 48);  // 8 x 6 floats = 12 x float4
        float a[48];
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const float4 t = a4[i];
            a[4 * i] = t.x; a[4 * i + 1] = t.y; a[4 * i + 2] = t.z; a[4 * i + 3] = t.w;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float2* x2 = reinterpret_cast<const float2*>(pb.p + 6 * (size_t) pb.nbr[8 * (size_t) v + k]);
            const float2 x01 = x2[0], x23 = x2[1], x45 = x2[2];
            acc = __fmaf_rn(a[6 * k], x01.x, acc); acc = __fmaf_rn(a[6 * k + 1], x01.y, acc);
            acc = __fmaf_rn(a[6 * k + 2], x23.x, acc); acc = __fmaf_rn(a[6 * k + 3], x23.y, acc);
            acc = __fmaf_rn(a[6 * k + 4], x45.x, acc); acc = __fmaf_rn(a[6 * k + 5], x45.y, acc);
        }
    }
    pb.sv[v] = th * acc;
}

__global__ void __launch_bounds__(TPB) kp_node_apply(P2PProblem pb, const Scalars* sc, int it) {
    if (it >= sc->done_it) return;
    __shared__ double sh[TPB / 32];
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    double pq = 0.0;
    for (int n = gw; n < pb.N; n += nw) {
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int j = pb.tptr[n] + lane; j < pb.tptr[n + 1]; j += 32) {
            const int v = pb.tv[j];
            const float s = pb.sv[v];
            const float2* a2 = reinterpret_cast<const float2*>(pb.jac + ((size_t) v * 8 + pb.tk[j]) * 6);
            const float2 a01 = a2[0], a23 = a2[1], a45 = a2[2];
            acc[0] = __fmaf_rn(a01.x, s, acc[0]); acc[1] = __fmaf_rn(a01.y, s, acc[1]);
            acc[2] = __fmaf_rn(a23.x, s, acc[2]); acc[3] = __fmaf_rn(a23.y, s, acc[3]);
            acc[4] = __fmaf_rn(a45.x, s, acc[4]); acc[5] = __fmaf_rn(a45.y, s, acc[5]);
        }
        if (pb.wreg2 > 0.f) p2p_reg_apply(pb, n, lane, pb.p, acc);
#pragma unroll
        for (int r = 0; r < 6; ++r) acc[r] = warp_sum(acc[r]);
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                pb.q[6 * (size_t) n + r] = acc[r];
                pq += (double) pb.p[6 * (size_t) n + r] * acc[r];
            }
        }
    }
    const double bs = block_sum(pq, sh);
    if (threadIdx.x == 0) pb.part[blockIdx.x] = bs;
}

__global__ void __launch_bounds__(TPB) kp_update(P2PProblem pb, const Scalars* sc, int it, int nblk_pq) {
    if (it >= sc->done_it) return;
    __shared__ double sh[TPB / 32];
    const double pq = sum_partials(pb.part, nblk_pq, sh);
    const double rz = sc->rz[it & 1];
    const float alpha = pq > 0.0 ? (float) (rz / pq) : 0.f;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    double rzn = 0.0;
    if (n < pb.N) {
        float r[6], z[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const size_t i = 6 * (size_t) n + c;
            pb.x[i] = __fmaf_rn(alpha, pb.p[i], pb.x[i]);
            r[c] = __fmaf_rn(-alpha, pb.q[i], pb.r[i]);
            pb.r[i] = r[c];
        }
        p2p_precond(pb.L + 21 * (size_t) n, r, z);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            pb.z[6 * (size_t) n + c] = z[c];
            rzn += (double) r[c] * z[c];
        }
    }
    const double bs = block_sum(rzn, sh);
    if (threadIdx.x == 0) pb.part[MAX_PARTIALS + blockIdx.x] = bs;
}

__global__ void __launch_bounds__(TPB) kp_direction(P2PProblem pb, Scalars* sc, int it, int nblk, int nblk_pq, double tol2) {
    if (it >= sc->done_it) return;
    __shared__ double sh[TPB / 32];
    const double rzn = sum_partials(pb.part + MAX_PARTIALS, nblk, sh);
    const double rz = sc->rz[it & 1];
    const float beta = rz > 0.0 ? (float) (rzn / rz) : 0.f;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < pb.N) {
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const size_t i = 6 * (size_t) n + c;
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

// X_n <- exp(xi_n) X_n  (double)
__global__ void __launch_bounds__(TPB) kp_expmap(P2PProblem pb, Scalars* sc) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 0) sc->gn_steps += 1;
    if (n >= pb.N) return;
    const float* xi = pb.x + 6 * (size_t) n;
    const double wx = xi[0], wy = xi[1], wz = xi[2];
    const double th2 = wx * wx + wy * wy + wz * wz, th = sqrt(th2);
    double A, B, C;
    if (th < 1e-6) {
        A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; C = 1.0 / 6.0 - th2 / 120.0;
    } else {
        A = sin(th) / th; B = (1.0 - cos(th)) / th2; C = (th - sin(th)) / (th2 * th);
    }
    const double K[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double K2[9], R[9], V[9], t[3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) K2[3 * r + c] = K[3 * r] * K[c] + K[3 * r + 1] * K[3 + c] + K[3 * r + 2] * K[6 + c];
    for (int i = 0; i < 9; ++i) {
        const double I = (i % 4 == 0) ? 1.0 : 0.0;
        R[i] = I + A * K[i] + B * K2[i];
        V[i] = I + B * K[i] + C * K2[i];
    }
    for (int r = 0; r < 3; ++r) t[r] = V[3 * r] * xi[3] + V[3 * r + 1] * xi[4] + V[3 * r + 2] * xi[5];
    float* X = pb.X + 12 * (size_t) n;
    double Xo[12];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) Xo[3 * r + c] = R[3 * r] * X[c] + R[3 * r + 1] * X[3 + c] + R[3 * r + 2] * X[6 + c];
        Xo[9 + r] = R[3 * r] * X[9] + R[3 * r + 1] * X[10] + R[3 * r + 2] * X[11] + t[r];
    }
    for (int i = 0; i < 12; ++i) X[i] = (float) Xo[i];
}

// dg_se3_n := DQ(X_n) * dg_se3_n  (DualQuaternion(rot, t), dual_quaternion.hpp:42-45; operator*, :127-129)
__global__ void __launch_bounds__(TPB) kp_compose(P2PProblem pb, float4* __restrict__ real, float4* __restrict__ dual) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= pb.N) return;
    const float* X = pb.X + 12 * (size_t) n;
    const double R[9] = {X[0], X[1], X[2], X[3], X[4], X[5], X[6], X[7], X[8]};
    double q[4];
    const double tr = R[0] + R[4] + R[8];
    if (tr > 0) {
        const double s = sqrt(tr + 1.0) * 2;
        q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        const double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
        q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s;
    } else if (R[4] > R[8]) {
        const double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
        q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s;
    } else {
        const double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
        q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s;
    }
    const Quat rot{(float) q[0], (float) q[1], (float) q[2], (float) q[3]};
    DQ inc;
    inc.real = qdiv(rot, qdot(rot, rot));  // boost's norm() is the squared norm (dual_quaternion.hpp:31,43)
    inc.dual = qscale(qmul(Quat{0.f, X[9], X[10], X[11]}, inc.real), 0.5f);
    const DQ cur{make_quat(real[n]), make_quat(dual[n])};
    const DQ out = dq_mul(inc, cur);
    real[n] = to_float4(out.real);
    dual[n] = to_float4(out.dual);
}
