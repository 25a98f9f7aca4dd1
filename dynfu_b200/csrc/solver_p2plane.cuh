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
//   atomics, reproducible; reductions in double.  Default: the whole solve in one cooperative launch (kp_persistent, 3 grid
//   barriers per PCG iteration); one kernel per phase when cooperative launch is unavailable or DFU_SOLVER_PATH=multi.
// (textually included by solver.cu inside its anonymous namespace -- one translation unit)
#pragma once

constexpr int P2P_VS = 8;  // floats per node in the PCG vectors

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
    float* Gd;     // N*8*3: their difference, formed in double
    float* ew;     // N*8: weight of out-edge (n, i): w_reg^2, times alpha_ij h_ij with the robust regulariser; 0 for self edges
    int reg_mode;  // DFU_REG_*
    float psi_reg;
    float *b, *x, *r, *z, *p, *q;  // N*P2P_VS (32-byte slot per node: one sector per gather)
    float* p2;     // second direction buffer of the persistent kernel
    // persistent kernel only
    float4* ent;   // 8P * 2: per transposed-graph entry (node <- point) {a0 a1 a2 a3 | a4 a5 theta e}: what the node gathers, in list order
    float* svT;    // 8P: theta (J x) of the entry's point, scattered by the point phase
    int* tpos;     // P*8: position of (point, slot) in the transposed lists
    float* Minv;   // N*36 (28 used): persistent kernel's preconditioner, packed Cholesky factor + reciprocal diagonal (all zero: singular block)
    unsigned char* rslot;  // 8N: for in-edge entry rin[i] = src of node n, the slot of n in src's out-edge list
    float* L;      // N*21: Cholesky factor of the diagonal block (row-major lower, packed), L[0] <= 0: singular block
    double* part;  // 4 * MAX_PARTIALS
};

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

// X_n c (X_n: 3x3 row-major | t, 48 bytes, 16-byte aligned: three 128-bit loads)
DFU_DEV void p2p_apply(const float* X, float cx, float cy, float cz, float& ox, float& oy, float& oz) {
    const float4* X4 = reinterpret_cast<const float4*>(X);
    const float4 r0 = X4[0], r1 = X4[1], r2 = X4[2];  // R00 R01 R02 R10 | R11 R12 R20 R21 | R22 tx ty tz
    ox = r0.x * cx + r0.y * cy + r0.z * cz + r2.y;
    oy = r0.w * cx + r1.x * cy + r1.y * cz + r2.z;
    oz = r1.z * cx + r1.w * cy + r2.x * cz + r2.w;
}

// the same in double: positions are O(1) m and residuals O(1) mm, so a float p_v - l_v would carry ~1e-7 of absolute noise
// into the gradient, which the flat directions of the energy amplify beyond the 1e-4 parity bound on the transforms
DFU_DEV void p2p_apply_d(const float* X, float cx, float cy, float cz, double& ox, double& oy, double& oz) {
    const float4* X4 = reinterpret_cast<const float4*>(X);
    const float4 r0 = X4[0], r1 = X4[1], r2 = X4[2];
    ox = (double) r0.x * cx + (double) r0.y * cy + (double) r0.z * cz + (double) r2.y;
    oy = (double) r0.w * cx + (double) r1.x * cy + (double) r1.y * cz + (double) r2.z;
    oz = (double) r1.z * cx + (double) r1.w * cy + (double) r2.x * cz + (double) r2.w;
}

// Persistent kernel: the point-major arrays it owns are stored "lane-contiguous" -- 12 arrays of float4 [P] for the Jacobians
// (element f = 6 k + c of point v lives in array f / 4) and 2 arrays of int4 [P] for tpos -- so a warp's load of one
// register's worth touches 4 cache lines instead of 32 (the point phase is bound by L1 wavefronts, not by bytes).
DFU_DEV float4* jac_s4(const P2PProblem& pb, int i, int v) { return reinterpret_cast<float4*>(pb.jac) + (size_t) i * pb.P + v; }
DFU_DEV float2* jac_s2(const P2PProblem& pb, int q, int v) {  // pair q = 3 k + c / 2 of point v
    return reinterpret_cast<float2*>(jac_s4(pb, q >> 1, v)) + (q & 1);
}
DFU_DEV int4* tpos_s4(const P2PProblem& pb, int h, int v) { return reinterpret_cast<int4*>(pb.tpos) + (size_t) h * pb.P + v; }
DFU_DEV int* tpos_s1(const P2PProblem& pb, int k, int v) { return reinterpret_cast<int*>(tpos_s4(pb, k >> 2, v)) + (k & 3); }

// linearisation point of point v: p_v, e_v = n.(p - l), the 8 Jacobian 6-vectors; optional Tukey update; returns theta e^2.
// SOA (persistent kernel): lane-contiguous Jacobians, and {a, theta, e} scattered to the point's 8 transposed entries (the
// node phases then read their lists contiguously instead of chasing point indices)
template <bool SOA>
DFU_DEV double p2p_linearise_point(const P2PProblem& pb, int v, bool update_tukey) {
    const float cx = pb.canon[3 * (size_t) v], cy = pb.canon[3 * (size_t) v + 1], cz = pb.canon[3 * (size_t) v + 2];
    const float nx = pb.nrm[3 * (size_t) v], ny = pb.nrm[3 * (size_t) v + 1], nz = pb.nrm[3 * (size_t) v + 2];
    const int4 n0 = *reinterpret_cast<const int4*>(pb.nbr + 8 * (size_t) v), n1 = *reinterpret_cast<const int4*>(pb.nbr + 8 * (size_t) v + 4);
    const int nb[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
    const float4 w0 = *reinterpret_cast<const float4*>(pb.wn + 8 * (size_t) v), w1 = *reinterpret_cast<const float4*>(pb.wn + 8 * (size_t) v + 4);
    const float wk[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    double px = 0.0, py = 0.0, pz = 0.0;
    float sw = 0.f;
    float a[48];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float w = wk[k];
        double qxd, qyd, qzd;
        p2p_apply_d(pb.X + 12 * (size_t) nb[k], cx, cy, cz, qxd, qyd, qzd);
        px += (double) w * qxd; py += (double) w * qyd; pz += (double) w * qzd;
        sw += w;
        const float qx = (float) qxd, qy = (float) qyd, qz = (float) qzd;
        a[6 * k] = w * (qy * nz - qz * ny); a[6 * k + 1] = w * (qz * nx - qx * nz); a[6 * k + 2] = w * (qx * ny - qy * nx);
        a[6 * k + 3] = w * nx; a[6 * k + 4] = w * ny; a[6 * k + 5] = w * nz;
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        float4* dst = SOA ? jac_s4(pb, i, v) : reinterpret_cast<float4*>(pb.jac + (size_t) v * 48) + i;
        *dst = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
    }
    const double dx = px - pb.live[3 * (size_t) v], dy = py - pb.live[3 * (size_t) v + 1], dz = pz - pb.live[3 * (size_t) v + 2];
    const double ed = (double) nx * dx + (double) ny * dy + (double) nz * dz;
    const float e = (float) ed;
    pb.e[v] = e;
    float th;
    if (update_tukey) {
        th = sw > 0.f ? tukey_biweight(pb.tukey_offset, pb.psi_data, -(float) dx, -(float) dy, -(float) dz) : 0.f;
        pb.theta[v] = th;
    } else {
        th = pb.theta[v];
    }
    if (SOA) {
        const int4 t0 = *tpos_s4(pb, 0, v), t1 = *tpos_s4(pb, 1, v);
        const int tp[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float4* r = pb.ent + 2 * (size_t) tp[k];
            r[0] = make_float4(a[6 * k], a[6 * k + 1], a[6 * k + 2], a[6 * k + 3]);
            r[1] = make_float4(a[6 * k + 4], a[6 * k + 5], th, e);
        }
    }
    return (double) th * ed * ed;
}
__global__ void __launch_bounds__(TPB) kp_linearise(P2PProblem pb, int update_tukey) {
    __shared__ double sh[TPB / 32];
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const double e2 = v < pb.P ? p2p_linearise_point<false>(pb, v, update_tukey != 0) : 0.0;
    const double bs = block_sum(e2, sh);
    if (threadIdx.x == 0) pb.part[blockIdx.x] = bs;
}

// X_n g_m and X_m g_m of out-edge i = n * 8 + k and their difference (formed in double); returns its regularisation energy.
// update_w: re-evaluate the edge weight -- w_reg^2, or with DFU_REG_HUBER_ALPHA (DynamicFusion eq. 8 as IRLS)
// w_reg^2 max(dg_w_n, dg_w_m) h with the Huber weight h = 1 (|d| <= psi_reg) or psi_reg / |d| (opt_solver.cpp:233-239) --
// which then stays frozen until the next outer iteration, like the Tukey weights of the data term
DFU_DEV double p2p_edge(const P2PProblem& pb, int i, bool update_w) {
    const int n = i >> 3, m = pb.nnbr[i];
    const float4 g = pb.pos_w[m];
    float* G = pb.G + 6 * (size_t) i;
    float* D = pb.Gd + 3 * (size_t) i;
    double a[3], b[3];
    p2p_apply_d(pb.X + 12 * (size_t) n, g.x, g.y, g.z, a[0], a[1], a[2]);
    p2p_apply_d(pb.X + 12 * (size_t) m, g.x, g.y, g.z, b[0], b[1], b[2]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        G[c] = (float) a[c];
        G[3 + c] = (float) b[c];
        D[c] = m == n ? 0.f : (float) (a[c] - b[c]);
    }
    const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2], d2 = dx * dx + dy * dy + dz * dz;
    float w;
    if (update_w) {
        double wd = m == n ? 0.0 : (double) pb.wreg2;
        if (pb.reg_mode == DFU_REG_HUBER_ALPHA) {
            const double r = sqrt(d2), psi = pb.psi_reg;
            wd *= (double) fmaxf(pb.pos_w[n].w, g.w) * (r <= psi ? 1.0 : psi / r);
        }
        w = (float) wd;
        pb.ew[i] = w;
    } else {
        w = pb.ew[i];
    }
    return m == n ? 0.0 : (double) w * d2;
}
__global__ void __launch_bounds__(TPB) kp_edges(P2PProblem pb, int update_w) {
    __shared__ double sh[TPB / 32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double er = i < pb.N * 8 ? p2p_edge(pb, i, update_w != 0) : 0.0;
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
DFU_DEV void p2p_reg_apply(const P2PProblem& pb, int n, int lane, const float* x, float (&acc)[6], int g = 32) {
    const int lo = pb.rin_ptr[n], hi = pb.rin_ptr[n + 1];
    for (int j = lane; j < 8 + 8 * (hi - lo); j += g) {
        // out-edges: (n, j); in-edges: every out-edge of a source `src` that points at n
        const bool out = j < 8;
        const int src = out ? n : pb.rin[lo + (j - 8) / 8];
        const int i = out ? j : (j - 8) & 7;
        const int m = pb.nnbr[(size_t) src * 8 + i];
        if (m == src || (!out && (m != n || src == n))) continue;
        const float* G = pb.G + 6 * ((size_t) src * 8 + i);
        const float* xs = x + P2P_VS * (size_t) src;
        const float* xm = x + P2P_VS * (size_t) m;
        const float r0 = (xs[1] * G[2] - xs[2] * G[1]) + xs[3] - (xm[1] * G[5] - xm[2] * G[4]) - xm[3];
        const float r1 = (xs[2] * G[0] - xs[0] * G[2]) + xs[4] - (xm[2] * G[3] - xm[0] * G[5]) - xm[4];
        const float r2 = (xs[0] * G[1] - xs[1] * G[0]) + xs[5] - (xm[0] * G[4] - xm[1] * G[3]) - xm[5];
        const float* Gk = out ? G : G + 3;
        const float we = pb.ew[(size_t) src * 8 + i];
        const float sg = out ? we : -we;
        acc[0] += sg * (Gk[1] * r2 - Gk[2] * r1);
        acc[1] += sg * (Gk[2] * r0 - Gk[0] * r2);
        acc[2] += sg * (Gk[0] * r1 - Gk[1] * r0);
        acc[3] += sg * r0; acc[4] += sg * r1; acc[5] += sg * r2;
    }
}

// regularisation part of the right-hand side and of the diagonal block of node n (edges touching n, g lanes per node)
DFU_DEV void p2p_assemble_reg(const P2PProblem& pb, int n, int lane, int g, double (&b)[6], double (&M)[21]) {
    const int lo = pb.rin_ptr[n], hi = pb.rin_ptr[n + 1];
    for (int j = lane; j < 8 + 8 * (hi - lo); j += g) {
        const bool out = j < 8;
        const int src = out ? n : pb.rin[lo + (j - 8) / 8];
        const int i = out ? j : (j - 8) & 7;
        const int m = pb.nnbr[(size_t) src * 8 + i];
        if (m == src || (!out && (m != n || src == n))) continue;
        const float* G = pb.G + 6 * ((size_t) src * 8 + i);
        const float* D = pb.Gd + 3 * ((size_t) src * 8 + i);
        const double r0 = D[0], r1 = D[1], r2 = D[2];
        const float* Gk = out ? G : G + 3;
        const double gx = Gk[0], gy = Gk[1], gz = Gk[2];
        const double w2 = pb.ew[(size_t) src * 8 + i], sg = out ? w2 : -w2;
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

// node n (one warp): b = -J^T r0, the 6x6 diagonal block and its Cholesky factor; PCG start x = 0, r = b, z = M^-1 b, p = z;
// returns the node's r.z on lane 0 (0 elsewhere)
DFU_DEV double p2p_assemble_node(const P2PProblem& pb, int n, int lane) {
    double rz = 0.0;
    {
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
        if (pb.wreg2 > 0.f) p2p_assemble_reg(pb, n, lane, 32, b, M);
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
                const size_t i = P2P_VS * (size_t) n + r;
                pb.b[i] = (float) b[r]; pb.r[i] = (float) b[r]; pb.x[i] = 0.f;
                pb.z[i] = (float) z[r]; pb.p[i] = (float) z[r];
                rz += b[r] * z[r];
            }
        }
    }
    return rz;
}
__global__ void __launch_bounds__(TPB) kp_assemble(P2PProblem pb) {
    __shared__ double sh[TPB / 32];
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    double rz = 0.0;
    for (int n = gw; n < pb.N; n += nw) rz += p2p_assemble_node(pb, n, lane);
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

// theta_v (J x)_v with x = p (FLY = false) or x = z + beta p, formed on the way (FLY = true: the persistent kernel folds
// the direction update into the product, one grid barrier less)
template <bool FLY>
DFU_DEV float p2p_point_dot(const P2PProblem& pb, int v, const float* p, const float* z, float beta) {
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
        const int4 n0 = *reinterpret_cast<const int4*>(pb.nbr + 8 * (size_t) v), n1 = *reinterpret_cast<const int4*>(pb.nbr + 8 * (size_t) v + 4);
        const int nb[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float2* x2 = reinterpret_cast<const float2*>(p + P2P_VS * (size_t) nb[k]);
            float2 x01 = x2[0], x23 = x2[1], x45 = x2[2];
            if (FLY) {
                const float2* z2 = reinterpret_cast<const float2*>(z + P2P_VS * (size_t) nb[k]);
                const float2 z01 = z2[0], z23 = z2[1], z45 = z2[2];
                x01.x = __fmaf_rn(beta, x01.x, z01.x); x01.y = __fmaf_rn(beta, x01.y, z01.y);
                x23.x = __fmaf_rn(beta, x23.x, z23.x); x23.y = __fmaf_rn(beta, x23.y, z23.y);
                x45.x = __fmaf_rn(beta, x45.x, z45.x); x45.y = __fmaf_rn(beta, x45.y, z45.y);
            }
            acc = __fmaf_rn(a[6 * k], x01.x, acc); acc = __fmaf_rn(a[6 * k + 1], x01.y, acc);
            acc = __fmaf_rn(a[6 * k + 2], x23.x, acc); acc = __fmaf_rn(a[6 * k + 3], x23.y, acc);
            acc = __fmaf_rn(a[6 * k + 4], x45.x, acc); acc = __fmaf_rn(a[6 * k + 5], x45.y, acc);
        }
    }
    return th * acc;
}
__global__ void __launch_bounds__(TPB) kp_point_apply(P2PProblem pb, const Scalars* sc, int it) {
    if (it >= sc->done_it) return;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= pb.P) return;
    pb.sv[v] = p2p_point_dot<false>(pb, v, pb.p, nullptr, 0.f);
}

// q_n = (J^T sv)_n + regularisation rows applied to x, by the warp of node n; lane 0 stores q and returns x_n . q_n
DFU_DEV double p2p_node_apply(const P2PProblem& pb, int n, int lane, const float* x) {
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
    if (pb.wreg2 > 0.f) p2p_reg_apply(pb, n, lane, x, acc);
#pragma unroll
    for (int r = 0; r < 6; ++r) acc[r] = warp_sum(acc[r]);
    double pq = 0.0;
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            pb.q[P2P_VS * (size_t) n + r] = acc[r];
            pq += (double) x[P2P_VS * (size_t) n + r] * acc[r];
        }
    }
    return pq;
}
__global__ void __launch_bounds__(TPB) kp_node_apply(P2PProblem pb, const Scalars* sc, int it) {
    if (it >= sc->done_it) return;
    __shared__ double sh[TPB / 32];
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    double pq = 0.0;
    for (int n = gw; n < pb.N; n += nw) pq += p2p_node_apply(pb, n, lane, pb.p);
    const double bs = block_sum(pq, sh);
    if (threadIdx.x == 0) pb.part[blockIdx.x] = bs;
}

// x_n += alpha p_n, r_n -= alpha q_n, z_n = M_n^-1 r_n; returns r_n . z_n
DFU_DEV double p2p_update_node(const P2PProblem& pb, int n, float alpha, const float* p) {
    float r[6], z[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const size_t i = P2P_VS * (size_t) n + c;
        pb.x[i] = __fmaf_rn(alpha, p[i], pb.x[i]);
        r[c] = __fmaf_rn(-alpha, pb.q[i], pb.r[i]);
        pb.r[i] = r[c];
    }
    p2p_precond(pb.L + 21 * (size_t) n, r, z);
    double rzn = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        pb.z[P2P_VS * (size_t) n + c] = z[c];
        rzn += (double) r[c] * z[c];
    }
    return rzn;
}
__global__ void __launch_bounds__(TPB) kp_update(P2PProblem pb, const Scalars* sc, int it, int nblk_pq) {
    if (it >= sc->done_it) return;
    __shared__ double sh[TPB / 32];
    const double pq = sum_partials(pb.part, nblk_pq, sh);
    const double rz = sc->rz[it & 1];
    const float alpha = pq > 0.0 ? (float) (rz / pq) : 0.f;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const double rzn = n < pb.N ? p2p_update_node(pb, n, alpha, pb.p) : 0.0;
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
            const size_t i = P2P_VS * (size_t) n + c;
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
DFU_DEV void p2p_expmap_node(const P2PProblem& pb, int n) {
    const float* xi = pb.x + P2P_VS * (size_t) n;
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
__global__ void __launch_bounds__(TPB) kp_expmap(P2PProblem pb, Scalars* sc) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 0) sc->gn_steps += 1;
    if (n < pb.N) p2p_expmap_node(pb, n);
}

// dg_se3_n := DQ(X_n) * dg_se3_n  (DualQuaternion(rot, t), dual_quaternion.hpp:42-45; operator*, :127-129)
DFU_DEV void p2p_compose_node(const P2PProblem& pb, int n, float4* __restrict__ real, float4* __restrict__ dual) {
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
__global__ void __launch_bounds__(TPB) kp_compose(P2PProblem pb, float4* __restrict__ real, float4* __restrict__ dual) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < pb.N) p2p_compose_node(pb, n, real, dual);
}

