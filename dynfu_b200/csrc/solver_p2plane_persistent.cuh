// Solver, north-star extension P2PLANE_SE3: the whole Gauss-Newton / PCG solve in one cooperative launch (kp_persistent).
// Uses the problem description and the per-item functions of solver_p2plane.cuh.
// (textually included by solver.cu inside its anonymous namespace -- one translation unit)
#pragma once

// ---------------------------------------------------------------------------------------------------------------------
// The whole solve in ONE cooperative launch: phases separated by grid barriers instead of kernel boundaries.
//   per GN step:       linearise + edges | assemble | (scalars, redundantly per CTA) ... PCG ... expmap |
//   per PCG iteration: sv = Theta J p, scattered to the transposed entries; owners store p = z + beta p_old |
//                      q = J^T sv + reg, p.q | alpha; x, r, z = M^-1 r, r.z |       -> 3 barriers (launch-per-phase: 4 launches)
// Layout differences from the launch-per-phase path, all to make the node phases stream instead of chase indices:
//   * every (point, slot) knows its position in the transposed lists (tpos); the linearisation writes {a, theta, e} there
//     (one 32-byte sector per entry) and the point phase scatters theta (J p) there, so a node reads its list contiguously;
//   * a node is served by g = 32 / 16 / 8 lanes (the largest g with N g <= threads: every node in one pass when possible);
//   * the block-Jacobi preconditioner is applied with the Cholesky factor and its reciprocal diagonal -- no divisions in the
//     PCG loop;
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
constexpr int P2P_FS = 28;  // floats per node of the stored preconditioner: 21 of the packed factor, 6 reciprocal diagonal entries, 1 pad
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
        const float we = pb.ew[(size_t) e.src * 8 + e.i];
        const float sg = e.out ? we : -we;
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
        const double w2 = pb.ew[ei], sg = e.out ? w2 : -w2;
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

// node n by its g lanes: b = -J^T r0 and the 6x6 diagonal block from the node's entry list (contiguous); lane 0 factors the
// block (double), stores the factor for the PCG loop and the PCG start (x = 0, r = b, z = M^-1 b, p = z); returns r.z there
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
    if (active && lig == 0) {
        double L[21], dinv[6], sol[6];
        const bool ok = p2p_cholesky(M, L, dinv);
        p2p_chol_solve(L, dinv, b, sol);
        // the preconditioner the PCG loop applies: the factor and its reciprocal diagonal in float (27 of the node's 28
        // floats; dinv[0] = 0 marks a block that is not positive definite: z = 0 there)
        float4* F = reinterpret_cast<float4*>(pb.Minv + P2P_FS * (size_t) n);
        float f[28];
#pragma unroll
        for (int i = 0; i < 21; ++i) f[i] = ok ? (float) L[i] : 0.f;
#pragma unroll
        for (int i = 0; i < 6; ++i) f[21 + i] = ok ? (float) dinv[i] : 0.f;
        f[27] = 0.f;
#pragma unroll
        for (int i = 0; i < 7; ++i) F[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        float bf[6], o[6];
        const float zero[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            bf[i] = (float) b[i];
            o[i] = ok ? (float) sol[i] : 0.f;
            rz += ok ? b[i] * sol[i] : 0.0;
        }
        const size_t s = P2P_VS * (size_t) n;
        store6(pb.b + s, bf); store6(pb.r + s, bf); store6(pb.x + s, zero); store6(pb.z + s, o); store6(pb.p + s, o);
    }
    return rz;
}

// point phase of one PCG iteration, one thread per point.  MODE 0: x = p (first iteration, p = z stored by the assembly);
// MODE 1: x = z and sv = theta (J z) + beta sv_old; MODE 2: x = p_old, the direction z + beta p_old formed on the way
template <int MODE>
DFU_DEV void p2p_point_phase(const P2PProblem& pb, int v, const float* x, const float* z, float beta) {
    // every load that does not depend on another one is issued up front (theta = 0 points are rare: no branch around them),
    // so the phase costs two memory round trips: {theta, Jacobians, neighbours, entry positions} and the gathers
    const float th = pb.theta[v];
    const int4 n0 = *reinterpret_cast<const int4*>(pb.nbr + 8 * (size_t) v), n1 = *reinterpret_cast<const int4*>(pb.nbr + 8 * (size_t) v + 4);
    const int4 t0 = *tpos_s4(pb, 0, v), t1 = *tpos_s4(pb, 1, v);
    const float sv_old = MODE == 1 ? pb.sv[v] : 0.f;
    float a[48];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const float4 t = *jac_s4(pb, i, v);
        a[4 * i] = t.x; a[4 * i + 1] = t.y; a[4 * i + 2] = t.z; a[4 * i + 3] = t.w;
    }
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
    float s = th * acc;
    if (MODE == 1) s = __fmaf_rn(beta, sv_old, s);
    if (th == 0.f) s = 0.f;  // (an infinite / NaN direction component must not leak through a zero weight)
    pb.sv[v] = s;
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

// q_n = (J^T sv)_n + regularisation rows applied to x, by the g lanes of node n; lane 0 stores q and returns x_n . q_n.
// (Lists are uneven -- at the headline size 3840 of 4096 nodes have points, 158 on average, 282 at most -- and a split into a
// data part over the nodes with points and an edge part over all nodes, each with its own lanes, was tried: no gain, the
// longest list still sets the pace and the second part adds its own latency chain.)
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

// x_n += alpha p_n, r_n -= alpha q_n, z_n = M_n^-1 r_n; returns r_n . z_n.  M_n^-1 is applied as two triangular solves with
// the stored Cholesky factor and its reciprocal diagonal -- no divisions, ~40 dependent multiply-adds.  (An explicit 6x6
// inverse in float is faster still but loses positive definiteness on the nearly singular blocks a locally planar patch
// produces -- three of a node's six directions are held by the regulariser only -- and the iteration then depends on the
// rounding of the inverse; (L L^T)^-1 is positive definite whatever the rounding of L.  Loading the node's state before
// alpha is known, across the CTA-wide sum of the p.q partials, was tried as well: the floats spill, slower.)
DFU_DEV double p2p_update_node_T(const P2PProblem& pb, int n, float alpha, const float* p) {
    const size_t s = P2P_VS * (size_t) n;
    float xv[6], rv[6], qv[6], pv[6], zv[6], f[28];
    load6(pb.x + s, xv); load6(pb.r + s, rv); load6(pb.q + s, qv); load6(p + s, pv);
    const float4* F = reinterpret_cast<const float4*>(pb.Minv + P2P_FS * (size_t) n);
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const float4 t = F[i];
        f[4 * i] = t.x; f[4 * i + 1] = t.y; f[4 * i + 2] = t.z; f[4 * i + 3] = t.w;
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        xv[c] = __fmaf_rn(alpha, pv[c], xv[c]);
        rv[c] = __fmaf_rn(-alpha, qv[c], rv[c]);
    }
    float y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        float t = rv[i];
#pragma unroll
        for (int k = 0; k < i; ++k) t = __fmaf_rn(-f[i * (i + 1) / 2 + k], y[k], t);
        y[i] = t * f[21 + i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        float t = y[i];
#pragma unroll
        for (int k = i + 1; k < 6; ++k) t = __fmaf_rn(-f[k * (k + 1) / 2 + i], zv[k], t);
        zv[i] = t * f[21 + i];
    }
    double rzn = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) rzn += (double) rv[i] * zv[i];
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
    for (int i = tid; i < pb.N * 8; i += T) er += p2p_edge(pb, i, update_tukey);
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
