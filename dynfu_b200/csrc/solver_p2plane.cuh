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
    float *b, *x, *r, *z, *p, *q;  // N*P2P_VS (32-byte slot per node: one sector per gather)
    float* p2;     // second direction buffer of the persistent kernel
    // persistent kernel only
    float4* ent;   // 8P * 2: per transposed-graph entry (node <- point) {a0 a1 a2 a3 | a4 a5 theta e}: what the node gathers, in list order
    float* svT;    // 8P: theta (J x) of the entry's point, scattered by the point phase
    int* tpos;     // P*8: position of (point, slot) in the transposed lists
    float* Minv;   // N*36: inverse of the diagonal block (all zero: singular block)
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

// X_n g_m and X_m g_m of out-edge i = n * 8 + k and their difference (formed in double); returns its regularisation energy
DFU_DEV double p2p_edge(const P2PProblem& pb, int i) {
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
    if (m == n) return 0.0;
    const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return (double) pb.wreg2 * (dx * dx + dy * dy + dz * dz);
}
__global__ void __launch_bounds__(TPB) kp_edges(P2PProblem pb) {
    __shared__ double sh[TPB / 32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double er = i < pb.N * 8 ? p2p_edge(pb, i) : 0.0;
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
        const float sg = out ? pb.wreg2 : -pb.wreg2;
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

// ---------------------------------------------------------------------------------------------------------------------
// The whole solve in ONE cooperative launch: phases separated by grid barriers instead of kernel boundaries.
//   per GN step:       linearise + edges | assemble | (scalars, redundantly per CTA) ... PCG ... expmap |
//   per PCG iteration: sv = Theta J p, scattered to the transposed entries; owners store p = z + beta p_old |
//                      q = J^T sv + reg, p.q | alpha; x, r, z = M^-1 r, r.z |       -> 3 barriers (launch-per-phase: 4 launches)
// Layout differences from the launch-per-phase path, all to make the node phases stream instead of chase indices:
//   * every (point, slot) knows its position in the transposed lists (tpos); the linearisation writes {a, theta, e} there
//     (one 32-byte sector per entry) and the point phase scatters theta (J p) there, so a node reads its list contiguously;
//   * a node is served by g = 32 / 16 / 8 lanes (the largest g with N g <= threads: every node in one pass when possible);
//   * the block-Jacobi preconditioner is applied as an explicit 6x6 inverse (computed in double from the Cholesky factor by
//     six lanes, one column each) -- no divisions in the PCG loop;
//   * the direction update is folded into the product: theta J (z + beta p_old) = theta J z + beta sv_old, so the point
//     phase gathers z only (recomputed from z and p_old every 16th iteration to stop rounding drift).
// Every CTA sums the per-CTA partials in the same fixed order, so all CTAs take the same decisions without a broadcast and
// the result does not depend on timing.
struct P2PCtl {
    int num_iter, nonlinear_iter, linear_iter, early_out;
    double tol2;
    int refresh;      // every refresh-th PCG iteration forms theta J p from z and p_old instead of the recurrence (1: always)
    long long* prof;  // DFU_SOLVER_PROFILE: SM cycles of CTA 0 per phase (P2P_PROF_N slots), else NULL
};
constexpr int P2P_TPB = 256;
constexpr int P2P_CTAS_PER_SM = 2;  // resident CTAs per SM the kernel is compiled for (128 registers; 3 CTAs / 80 registers spill in the point phase: 1.60 vs 1.28 ms)
constexpr int P2P_PROF_N = 16;
// phase timer of CTA 0 (thread 0): adds the cycles since the previous mark to slot i
#define P2P_MARK(i)                                        \
    if (ctl.prof != nullptr && tid == 0) {                 \
        const long long t_now = clock64();                 \
        ctl.prof[i] += t_now - t_mark;                     \
        t_mark = t_now;                                    \
    }

// sum over the g lanes serving one node (xor butterfly: every lane ends with the total).  All 32 lanes execute all five
// steps -- the groups of one warp may have different sizes -- and a lane adds only the steps inside its own group.
DFU_DEV double group_sum(double v, int g) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double t = __shfl_xor_sync(0xffffffffu, v, o);
        if (o < g) v += t;
    }
    return v;
}
DFU_DEV float group_sum(float v, int g) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float t = __shfl_xor_sync(0xffffffffu, v, o);
        if (o < g) v += t;
    }
    return v;
}
DFU_DEV void load6(const float* s, float (&v)[6]) {  // 32-byte aligned slot
    const float4 a = *reinterpret_cast<const float4*>(s);
    const float2 b = *reinterpret_cast<const float2*>(s + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y;
}
DFU_DEV void store6(float* d, const float (&v)[6]) {
    *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float2*>(d + 4) = make_float2(v[4], v[5]);
}

// Cholesky factor of a packed lower-triangular 6x6 block, fully unrolled (registers); dinv = 1 / diagonal
DFU_DEV bool p2p_cholesky(const double (&M)[21], double (&L)[21], double (&dinv)[6]) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double s = M[i * (i + 1) / 2 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
            if (i == j) {
                ok = ok && (s > 0.0);
                const double d = sqrt(ok ? s : 1.0);
                L[i * (i + 1) / 2 + j] = d;
                dinv[i] = 1.0 / d;
            } else {
                L[i * (i + 1) / 2 + j] = s * dinv[j];
            }
        }
    }
    return ok;
}
DFU_DEV void p2p_chol_solve(const double (&L)[21], const double (&dinv)[6], const double (&rhs)[6], double (&out)[6]) {
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double s = rhs[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s -= L[i * (i + 1) / 2 + k] * y[k];
        y[i] = s * dinv[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double s = y[i];
#pragma unroll
        for (int k = i + 1; k < 6; ++k) s -= L[k * (k + 1) / 2 + i] * out[k];
        out[i] = s * dinv[i];
    }
}

// The edges touching node n, for the persistent kernel: its 8 out-edges and its in-edges, one candidate per edge (rslot
// names the slot of n in the source's list; the launch-per-phase path scans all 8 slots of every in-neighbour instead).
// Self edges carry no residual.
struct P2PEdge {
    bool live, out;
    int src, m, i;
};
DFU_DEV P2PEdge p2p_edge_of(const P2PProblem& pb, int n, int lo, int j) {
    P2PEdge e;
    e.out = j < 8;
    e.src = e.out ? n : pb.rin[lo + j - 8];
    e.i = e.out ? j : pb.rslot[lo + j - 8];
    e.m = e.out ? pb.nnbr[(size_t) n * 8 + j] : n;
    e.live = e.m != e.src;
    return e;
}
DFU_DEV void p2p_reg_apply_T(const P2PProblem& pb, int n, int lig, int g, const float* x, float (&acc)[6]) {
    const int lo = pb.rin_ptr[n], cnt = 8 + pb.rin_ptr[n + 1] - lo;
    for (int j = lig; j < cnt; j += g) {
        const P2PEdge e = p2p_edge_of(pb, n, lo, j);
        if (!e.live) continue;
        const float2* G2 = reinterpret_cast<const float2*>(pb.G + 6 * ((size_t) e.src * 8 + e.i));
        const float2 g01 = G2[0], g23 = G2[1], g45 = G2[2];
        const float G[6] = {g01.x, g01.y, g23.x, g23.y, g45.x, g45.y};
        float xs[6], xm[6];
        load6(x + P2P_VS * (size_t) e.src, xs);
        load6(x + P2P_VS * (size_t) e.m, xm);
        const float r0 = (xs[1] * G[2] - xs[2] * G[1]) + xs[3] - (xm[1] * G[5] - xm[2] * G[4]) - xm[3];
        const float r1 = (xs[2] * G[0] - xs[0] * G[2]) + xs[4] - (xm[2] * G[3] - xm[0] * G[5]) - xm[4];
        const float r2 = (xs[0] * G[1] - xs[1] * G[0]) + xs[5] - (xm[0] * G[4] - xm[1] * G[3]) - xm[5];
        const float* Gk = e.out ? G : G + 3;
        const float sg = e.out ? pb.wreg2 : -pb.wreg2;
        acc[0] += sg * (Gk[1] * r2 - Gk[2] * r1);
        acc[1] += sg * (Gk[2] * r0 - Gk[0] * r2);
        acc[2] += sg * (Gk[0] * r1 - Gk[1] * r0);
        acc[3] += sg * r0; acc[4] += sg * r1; acc[5] += sg * r2;
    }
}
DFU_DEV void p2p_assemble_reg_T(const P2PProblem& pb, int n, int lig, int g, double (&b)[6], double (&M)[21]) {
    const int lo = pb.rin_ptr[n], cnt = 8 + pb.rin_ptr[n + 1] - lo;
    for (int j = lig; j < cnt; j += g) {
        const P2PEdge e = p2p_edge_of(pb, n, lo, j);
        if (!e.live) continue;
        const size_t ei = (size_t) e.src * 8 + e.i;
        const float* Gk = pb.G + 6 * ei + (e.out ? 0 : 3);
        const float* D = pb.Gd + 3 * ei;
        const double r0 = D[0], r1 = D[1], r2 = D[2];
        const double gx = Gk[0], gy = Gk[1], gz = Gk[2];
        const double sg = e.out ? (double) pb.wreg2 : -(double) pb.wreg2, w2 = pb.wreg2;
        b[0] -= sg * (gy * r2 - gz * r1); b[1] -= sg * (gz * r0 - gx * r2); b[2] -= sg * (gx * r1 - gy * r0);
        b[3] -= sg * r0; b[4] -= sg * r1; b[5] -= sg * r2;
        M[0] += w2 * (gy * gy + gz * gz);
        M[1] += w2 * (-gx * gy); M[2] += w2 * (gx * gx + gz * gz);
        M[3] += w2 * (-gx * gz); M[4] += w2 * (-gy * gz); M[5] += w2 * (gx * gx + gy * gy);
        M[7] += w2 * gz;   M[8] += w2 * (-gy);  M[9] += w2;
        M[10] += w2 * (-gz); M[12] += w2 * gx;    M[14] += w2;
        M[15] += w2 * gy;   M[16] += w2 * (-gx); M[20] += w2;
    }
}

// node n by its g lanes: b = -J^T r0 and the 6x6 diagonal block from the node's entry list (contiguous), then up to seven
// lanes solve for the six columns of the inverse and for the PCG start (x = 0, r = b, z = M^-1 b, p = z); returns r.z there
DFU_DEV double p2p_assemble_node_T(const P2PProblem& pb, int n, bool active, int lig, int g) {
    double b[6] = {0, 0, 0, 0, 0, 0}, M[21];
#pragma unroll
    for (int i = 0; i < 21; ++i) M[i] = 0.0;
    if (active) {
        for (int j = pb.tptr[n] + lig; j < pb.tptr[n + 1]; j += g) {
            const float4 A = pb.ent[2 * (size_t) j], B = pb.ent[2 * (size_t) j + 1];
            const float a[6] = {A.x, A.y, A.z, A.w, B.x, B.y};
            const double th = B.z, te = th * (double) B.w;
            int idx = 0;
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                b[r] -= te * a[r];
#pragma unroll
                for (int c = 0; c <= r; ++c) M[idx++] += th * (double) a[r] * (double) a[c];
            }
        }
        if (pb.wreg2 > 0.f) p2p_assemble_reg_T(pb, n, lig, g, b, M);
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) b[r] = group_sum(b[r], g);
#pragma unroll
    for (int i = 0; i < 21; ++i) M[i] = group_sum(M[i], g);
    double rz = 0.0;
    if (active && lig < 7) {
        double L[21], dinv[6];
        const bool ok = p2p_cholesky(M, L, dinv);
        // seven solves shared by the group's lanes: roles 0..5 = columns of the inverse, role 6 = the PCG start
        for (int role = lig; role < 7; role += g) {
            double rhs[6], sol[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) rhs[i] = role < 6 ? (i == role ? 1.0 : 0.0) : b[i];
            p2p_chol_solve(L, dinv, rhs, sol);
            float o[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) o[i] = ok ? (float) sol[i] : 0.f;
            if (role < 6) {  // column `role` of the (symmetric) inverse
                float* Mi = pb.Minv + 36 * (size_t) n + 6 * role;
#pragma unroll
                for (int i = 0; i < 6; ++i) Mi[i] = o[i];
            } else {
                float bf[6];
                const float zero[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    bf[i] = (float) b[i];
                    rz += ok ? b[i] * sol[i] : 0.0;
                }
                const size_t s = P2P_VS * (size_t) n;
                store6(pb.b + s, bf); store6(pb.r + s, bf); store6(pb.x + s, zero); store6(pb.z + s, o); store6(pb.p + s, o);
            }
        }
    }
    return rz;
}

// point phase of one PCG iteration, one thread per point.  MODE 0: x = p (first iteration, p = z stored by the assembly);
// MODE 1: x = z and sv = theta (J z) + beta sv_old; MODE 2: x = p_old, the direction z + beta p_old formed on the way
template <int MODE>
DFU_DEV void p2p_point_phase(const P2PProblem& pb, int v, const float* x, const float* z, float beta) {
    const float th = pb.theta[v];
    float s = 0.f;
    if (th != 0.f) {
        float a[48];
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const float4 t = *jac_s4(pb, i, v);
            a[4 * i] = t.x; a[4 * i + 1] = t.y; a[4 * i + 2] = t.z; a[4 * i + 3] = t.w;
        }
        const int4 n0 = *reinterpret_cast<const int4*>(pb.nbr + 8 * (size_t) v), n1 = *reinterpret_cast<const int4*>(pb.nbr + 8 * (size_t) v + 4);
        const int nb[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float xv[6];
            load6(x + P2P_VS * (size_t) nb[k], xv);
            if (MODE == 2) {
                float zv[6];
                load6(z + P2P_VS * (size_t) nb[k], zv);
#pragma unroll
                for (int c = 0; c < 6; ++c) xv[c] = __fmaf_rn(beta, xv[c], zv[c]);
            }
#pragma unroll
            for (int c = 0; c < 6; ++c) acc = __fmaf_rn(a[6 * k + c], xv[c], acc);
        }
        s = th * acc;
        if (MODE == 1) s = __fmaf_rn(beta, pb.sv[v], s);
    }
    pb.sv[v] = s;
    const int4 t0 = *tpos_s4(pb, 0, v), t1 = *tpos_s4(pb, 1, v);
    pb.svT[t0.x] = s; pb.svT[t0.y] = s; pb.svT[t0.z] = s; pb.svT[t0.w] = s;
    pb.svT[t1.x] = s; pb.svT[t1.y] = s; pb.svT[t1.z] = s; pb.svT[t1.w] = s;
}

// sum over the 8 lanes that share a point in the point phases (lane = slot k of the point)
DFU_DEV double slot_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}
DFU_DEV float slot_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// the same with one lane per (point, slot) e = 8 v + k: more instructions per point, but an eighth of the latency -- used
// for the points left over after the full rounds of the per-point mapping (P mod T), so that no thread serves two points
// while the others wait at the barrier
template <int MODE>
DFU_DEV void p2p_point_phase_slot(const P2PProblem& pb, long e, bool valid, const float* x, const float* z, float beta) {
    float acc = 0.f, th = 0.f;
    const int v = (int) (e >> 3);
    if (valid) {
        th = pb.theta[v];
        if (th != 0.f) {
            const int k3 = 3 * (int) (e & 7);
            const float2 a01 = *jac_s2(pb, k3, v), a23 = *jac_s2(pb, k3 + 1, v), a45 = *jac_s2(pb, k3 + 2, v);
            const size_t o = P2P_VS * (size_t) pb.nbr[e];
            float xv[6];
            load6(x + o, xv);
            if (MODE == 2) {
                float zv[6];
                load6(z + o, zv);
#pragma unroll
                for (int c = 0; c < 6; ++c) xv[c] = __fmaf_rn(beta, xv[c], zv[c]);
            }
            acc = a01.x * xv[0];
            acc = __fmaf_rn(a01.y, xv[1], acc); acc = __fmaf_rn(a23.x, xv[2], acc); acc = __fmaf_rn(a23.y, xv[3], acc);
            acc = __fmaf_rn(a45.x, xv[4], acc); acc = __fmaf_rn(a45.y, xv[5], acc);
        }
    }
    acc = slot_sum(acc);
    const bool head = valid && (e & 7) == 0;
    float s = th * acc;
    if (MODE == 1) {  // the point's first lane owns sv_v
        if (head) s = __fmaf_rn(beta, pb.sv[v], s);
        s = __shfl_sync(0xffffffffu, s, (threadIdx.x & 31) & ~7);
    }
    if (head) pb.sv[v] = s;
    if (valid) pb.svT[*tpos_s1(pb, (int) (e & 7), v)] = s;
}

template <int MODE>
DFU_DEV void p2p_phase_points(const P2PProblem& pb, int tid, int T, const float* x, const float* z, float beta) {
    const int full = (pb.P / T) * T;  // points served one per thread
    for (int v = tid; v < full; v += T) p2p_point_phase<MODE>(pb, v, x, z, beta);
    const long e0 = 8L * full, e1 = 8L * pb.P;
    // remainder: its 8 (P - full) lanes are taken from the END of the grid, away from CTA 0 which sums up last
    for (long base = e0; base < e1; base += T) {
        const long e = base + (T - 1 - tid) / 8 * 8 + (tid & 7);
        p2p_point_phase_slot<MODE>(pb, e, e < e1, x, z, beta);
    }
}

// q_n = (J^T sv)_n + regularisation rows applied to x, by the g lanes of node n; lane 0 stores q and returns x_n . q_n
DFU_DEV double p2p_node_apply_T(const P2PProblem& pb, int n, bool active, int lig, int g, const float* x) {
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (active) {
        const int hi = pb.tptr[n + 1];
#pragma unroll 4
        for (int j = pb.tptr[n] + lig; j < hi; j += g) {
            const float4 A = pb.ent[2 * (size_t) j];
            const float2 B = *reinterpret_cast<const float2*>(pb.ent + 2 * (size_t) j + 1);
            const float s = pb.svT[j];
            acc[0] = __fmaf_rn(A.x, s, acc[0]); acc[1] = __fmaf_rn(A.y, s, acc[1]);
            acc[2] = __fmaf_rn(A.z, s, acc[2]); acc[3] = __fmaf_rn(A.w, s, acc[3]);
            acc[4] = __fmaf_rn(B.x, s, acc[4]); acc[5] = __fmaf_rn(B.y, s, acc[5]);
        }
        if (pb.wreg2 > 0.f) p2p_reg_apply_T(pb, n, lig, g, x, acc);
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) acc[r] = group_sum(acc[r], g);
    double pq = 0.0;
    if (active && lig == 0) {
        float xv[6];
        load6(x + P2P_VS * (size_t) n, xv);
        store6(pb.q + P2P_VS * (size_t) n, acc);
#pragma unroll
        for (int r = 0; r < 6; ++r) pq += (double) xv[r] * acc[r];
    }
    return pq;
}

// x_n += alpha p_n, r_n -= alpha q_n, z_n = M_n^-1 r_n (explicit inverse); returns r_n . z_n
DFU_DEV double p2p_update_node_T(const P2PProblem& pb, int n, float alpha, const float* p) {
    const size_t s = P2P_VS * (size_t) n;
    float xv[6], rv[6], qv[6], pv[6], zv[6], Mi[36];
    load6(pb.x + s, xv); load6(pb.r + s, rv); load6(pb.q + s, qv); load6(p + s, pv);
    const float4* M4 = reinterpret_cast<const float4*>(pb.Minv + 36 * (size_t) n);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const float4 t = M4[i];
        Mi[4 * i] = t.x; Mi[4 * i + 1] = t.y; Mi[4 * i + 2] = t.z; Mi[4 * i + 3] = t.w;
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        xv[c] = __fmaf_rn(alpha, pv[c], xv[c]);
        rv[c] = __fmaf_rn(-alpha, qv[c], rv[c]);
    }
    double rzn = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) t = __fmaf_rn(Mi[6 * i + j], rv[j], t);
        zv[i] = t;
        rzn += (double) rv[i] * t;
    }
    store6(pb.x + s, xv); store6(pb.r + s, rv); store6(pb.z + s, zv);
    return rzn;
}

// linearisation, one lane per (point, slot): the lane's Jacobian 6-vector, the point's position / residual / Tukey weight
// by a sum over its 8 lanes; {a, theta, e} goes to the transposed entry, a also to the point-major copy.  Returns theta e^2
// on the point's first lane.
DFU_DEV double p2p_linearise_slot(const P2PProblem& pb, long e, bool valid, bool update_tukey) {
    const int v = (int) (e >> 3), lane = threadIdx.x & 31;
    double px = 0.0, py = 0.0, pz = 0.0, e2 = 0.0;
    float sw = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
    float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (valid) {
        const float cx = pb.canon[3 * (size_t) v], cy = pb.canon[3 * (size_t) v + 1], cz = pb.canon[3 * (size_t) v + 2];
        nx = pb.nrm[3 * (size_t) v]; ny = pb.nrm[3 * (size_t) v + 1]; nz = pb.nrm[3 * (size_t) v + 2];
        const float w = pb.wn[e];
        double qxd, qyd, qzd;
        p2p_apply_d(pb.X + 12 * (size_t) pb.nbr[e], cx, cy, cz, qxd, qyd, qzd);
        px = (double) w * qxd; py = (double) w * qyd; pz = (double) w * qzd;
        sw = w;
        const float qx = (float) qxd, qy = (float) qyd, qz = (float) qzd;
        a[0] = w * (qy * nz - qz * ny); a[1] = w * (qz * nx - qx * nz); a[2] = w * (qx * ny - qy * nx);
        a[3] = w * nx; a[4] = w * ny; a[5] = w * nz;
        const int k3 = 3 * (int) (e & 7);
        *jac_s2(pb, k3, v) = make_float2(a[0], a[1]); *jac_s2(pb, k3 + 1, v) = make_float2(a[2], a[3]);
        *jac_s2(pb, k3 + 2, v) = make_float2(a[4], a[5]);
    }
    px = slot_sum(px); py = slot_sum(py); pz = slot_sum(pz); sw = slot_sum(sw);
    float ev = 0.f, th = 0.f;
    const bool head = valid && (e & 7) == 0;
    if (head) {
        const double dx = px - pb.live[3 * (size_t) v], dy = py - pb.live[3 * (size_t) v + 1], dz = pz - pb.live[3 * (size_t) v + 2];
        const double ed = (double) nx * dx + (double) ny * dy + (double) nz * dz;
        ev = (float) ed;
        pb.e[v] = ev;
        if (update_tukey) {
            th = sw > 0.f ? tukey_biweight(pb.tukey_offset, pb.psi_data, -(float) dx, -(float) dy, -(float) dz) : 0.f;
            pb.theta[v] = th;
        } else {
            th = pb.theta[v];
        }
        e2 = (double) th * ed * ed;
    }
    ev = __shfl_sync(0xffffffffu, ev, lane & ~7);
    th = __shfl_sync(0xffffffffu, th, lane & ~7);
    if (valid) {
        float4* r = pb.ent + 2 * (size_t) *tpos_s1(pb, (int) (e & 7), v);
        r[0] = make_float4(a[0], a[1], a[2], a[3]);
        r[1] = make_float4(a[4], a[5], th, ev);
    }
    return e2;
}

DFU_DEV double p2p_phase_linearise(const P2PProblem& pb, bool update_tukey, int tid, int T, double& er) {
    double ed = 0.0;
    const int full = (pb.P / T) * T;
    for (int v = tid; v < full; v += T) ed += p2p_linearise_point<true>(pb, v, update_tukey);
    const long e0 = 8L * full, e1 = 8L * pb.P;
    for (long base = e0; base < e1; base += T) {
        const long e = base + (T - 1 - tid) / 8 * 8 + (tid & 7);
        ed += p2p_linearise_slot(pb, e, e < e1, update_tukey);
    }
    for (int i = tid; i < pb.N * 8; i += T) er += p2p_edge(pb, i);
    return ed;
}

__global__ void __launch_bounds__(P2P_TPB, P2P_CTAS_PER_SM)
kp_persistent(P2PProblem pb, P2PCtl ctl, Scalars* sc, unsigned* bar, float4* __restrict__ real, float4* __restrict__ dual) {
    __shared__ double sh[4 * (P2P_TPB / 32)];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, T = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const int nb = gridDim.x;
    // lanes per node in the node phases, the node of this lane's group in pass `base`: base + grp
    const int g = (long) pb.N * 32 <= T ? 32 : ((long) pb.N * 16 <= T ? 16 : 8);
    const int lig = lane & (g - 1), grp = tid / g, ngrp = T / g;
    const int n0 = blockIdx.x + nb * threadIdx.x;  // first node of this thread in the thread-per-node phases (spread over the SMs)
    unsigned target = 0;
    double* part_ed = pb.part;
    double* part_er = pb.part + MAX_PARTIALS;
    double* part_rz = pb.part + 2 * MAX_PARTIALS;
    double* part_pq = pb.part + 3 * MAX_PARTIALS;
    float* pbuf[2] = {pb.p, pb.p2};
    long long t_mark = clock64();

    // ---- initial state: X = identity, normalised weights, position of every (point, slot) in the transposed lists
    for (int n = n0; n < pb.N; n += T) {
#pragma unroll
        for (int k = 0; k < 12; ++k) pb.X[12 * (size_t) n + k] = (k == 0 || k == 4 || k == 8) ? 1.f : 0.f;
    }
    for (int v = tid; v < pb.P; v += T) {
        float w[8], s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            w[k] = pb.wts[8 * (size_t) v + k];
            s += w[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) pb.wn[8 * (size_t) v + k] = s > 0.f ? w[k] / s : 0.f;
    }
    for (int n = tid >> 5; n < pb.N; n += T >> 5) {  // (one warp per node)
        for (int j = pb.tptr[n] + (tid & 31); j < pb.tptr[n + 1]; j += 32) {
            const int v = pb.tv[j];
            *tpos_s1(pb, p2p_slot(pb, v, n), v) = j;
        }
        for (int j = pb.rin_ptr[n] + (tid & 31); j < pb.rin_ptr[n + 1]; j += 32) {
            const int src = pb.rin[j];
            int k = 0;
#pragma unroll
            for (int i = 1; i < 8; ++i) k = pb.nnbr[(size_t) src * 8 + i] == n ? i : k;
            pb.rslot[j] = (unsigned char) k;
        }
    }
    grid_barrier(bar, nb, target);
    P2P_MARK(0);

    double E = 0.0, E0 = 0.0, rz_ref = -1.0;
    bool first = true, stop_all = false;
    int pcg_iters = 0, gn_steps = 0;
    for (int outer = 0; outer < ctl.num_iter && !stop_all; ++outer) {
        for (int gn = 0; gn < ctl.nonlinear_iter; ++gn) {
            // ---- linearisation point, edge transforms
            {
                double er = 0.0;
                const double ed = p2p_phase_linearise(pb, gn == 0, tid, T, er);
                const D4 s4 = block_sum4(D4{ed, er, 0.0, 0.0}, sh);
                if (threadIdx.x == 0) {
                    part_ed[blockIdx.x] = s4.a;
                    part_er[blockIdx.x] = s4.b;
                }
            }
            P2P_MARK(1);
            grid_barrier(bar, nb, target);
            P2P_MARK(2);
            // ---- per-node blocks, right-hand side, PCG start
            {
                double rzp = 0.0;
                for (int base = 0; base < pb.N; base += ngrp) rzp += p2p_assemble_node_T(pb, base + grp, base + grp < pb.N, lig, g);
                const double bs = block_sum(rzp, sh);
                if (threadIdx.x == 0) part_rz[blockIdx.x] = bs;
            }
            P2P_MARK(3);
            grid_barrier(bar, nb, target);
            P2P_MARK(4);
            double rz;
            {
                D4 v{0.0, 0.0, 0.0, 0.0};
                for (int i = threadIdx.x; i < nb; i += blockDim.x) {
                    v.a += part_ed[i]; v.b += part_er[i]; v.c += part_rz[i];
                }
                v = block_sum4(v, sh);
                E = v.a + v.b;
                rz = v.c;
            }
            if (first) {
                E0 = E;
                first = false;
            }
            if (rz_ref < 0.0) rz_ref = rz;
            P2P_MARK(5);
            bool done = !(rz > 0.0) || rz <= ctl.tol2 * rz_ref;
            if (done && ctl.early_out) {  // the launch-per-phase path reads the scalars back here and leaves the loops
                if (gn == 0 && outer > 0) stop_all = true;
                // the partials of this evaluation are overwritten by the next one: every CTA must have read them
                grid_barrier(bar, nb, target);
                break;
            }
            // ---- PCG
            float beta = 0.f;
            for (int it = 0; it < ctl.linear_iter && !done; ++it) {
                float* pn = pbuf[it & 1];
                const float* po = pbuf[(it & 1) ^ 1];
                if (it == 0) {
                    p2p_phase_points<0>(pb, tid, T, pn, nullptr, 0.f);
                } else {
                    if (it % ctl.refresh) {
                        p2p_phase_points<1>(pb, tid, T, pb.z, nullptr, beta);
                    } else {
                        p2p_phase_points<2>(pb, tid, T, po, pb.z, beta);
                    }
                    if (lig == 0)
                        for (int n = grp; n < pb.N; n += ngrp) {  // the node's owner stores the new direction
                            const size_t s = P2P_VS * (size_t) n;
                            float zv[6], pv[6];
                            load6(pb.z + s, zv); load6(po + s, pv);
#pragma unroll
                            for (int c = 0; c < 6; ++c) pv[c] = __fmaf_rn(beta, pv[c], zv[c]);
                            store6(pn + s, pv);
                        }
                }
                P2P_MARK(6);
                grid_barrier(bar, nb, target);
                P2P_MARK(7);
                {
                    double pqp = 0.0;
                    for (int base = 0; base < pb.N; base += ngrp) pqp += p2p_node_apply_T(pb, base + grp, base + grp < pb.N, lig, g, pn);
                    const double bs = block_sum(pqp, sh);
                    if (threadIdx.x == 0) part_pq[blockIdx.x] = bs;
                }
                P2P_MARK(8);
                grid_barrier(bar, nb, target);
                P2P_MARK(9);
                double pq;
                {
                    double v = 0.0;
                    for (int i = threadIdx.x; i < nb; i += blockDim.x) v += part_pq[i];
                    pq = block_sum(v, sh);
                }
                const float alpha = pq > 0.0 ? (float) (rz / pq) : 0.f;
                {
                    double rzp = 0.0;
                    if (lig == 0)
                        for (int n = grp; n < pb.N; n += ngrp) rzp += p2p_update_node_T(pb, n, alpha, pn);
                    const double bs = block_sum(rzp, sh);
                    if (threadIdx.x == 0) part_rz[blockIdx.x] = bs;
                }
                P2P_MARK(10);
                grid_barrier(bar, nb, target);
                P2P_MARK(11);
                double rzn;
                {
                    double v = 0.0;
                    for (int i = threadIdx.x; i < nb; i += blockDim.x) v += part_rz[i];
                    rzn = block_sum(v, sh);
                }
                beta = rz > 0.0 ? (float) (rzn / rz) : 0.f;
                rz = rzn;
                pcg_iters += 1;
                if (!(pq > 0.0) || !(rzn > 0.0) || rzn <= ctl.tol2 * rz_ref) done = true;
                P2P_MARK(12);
            }
            // ---- X <- exp(xi) X
            for (int n = n0; n < pb.N; n += T) p2p_expmap_node(pb, n);
            gn_steps += 1;
            grid_barrier(bar, nb, target);
            P2P_MARK(13);
        }
    }
    // ---- energy at the solution (Tukey weights of the last outer iteration; at the identity if no step ran)
    {
        double er = 0.0;
        const double ed = p2p_phase_linearise(pb, ctl.num_iter * ctl.nonlinear_iter == 0, tid, T, er);
        const D4 s4 = block_sum4(D4{ed, er, 0.0, 0.0}, sh);
        if (threadIdx.x == 0) {
            part_ed[blockIdx.x] = s4.a;
            part_er[blockIdx.x] = s4.b;
        }
    }
    grid_barrier(bar, nb, target);
    if (blockIdx.x == 0) {
        D4 v{0.0, 0.0, 0.0, 0.0};
        for (int i = threadIdx.x; i < nb; i += blockDim.x) {
            v.a += part_ed[i]; v.b += part_er[i];
        }
        v = block_sum4(v, sh);
        if (threadIdx.x == 0) {
            E = v.a + v.b;
            sc->E = E;
            sc->E0 = first ? E : E0;
            sc->first = 0;
            sc->rz[0] = sc->rz[1] = 0.0;
            sc->rz_ref = rz_ref;
            sc->done_it = INT_MAX;
            sc->pcg_iters = pcg_iters;
            sc->gn_steps = gn_steps;
            sc->spin_fail = 0;
        }
    }
    // ---- compose the increments onto the nodes once, like the reference does with its translations (opt_solver.cpp:270-285)
    for (int n = n0; n < pb.N; n += T) p2p_compose_node(pb, n, real, dual);
    P2P_MARK(14);
}
