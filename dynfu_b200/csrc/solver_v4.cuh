// Solver: persistent cooperative kernel, version 4 (default for the frame loop)
// (textually included by solver.cu inside its anonymous namespace -- one translation unit)
#pragma once

// =====================================================================================================
// Version 4 = version 3r (explicit normal matrix, rows in registers, pipelined PCG of Ghysels & Vanroose) with the two
// things that bounded 3r removed (profiles/r02_solver_v4.md; measured on B200 with DFU_SOLVER_PROFILE=1):
//
//  (1) NO grid barrier inside the PCG loop.  3r spent 2.5 k cycles per iteration in the barrier and 2.8 k in the dependent
//      loads behind it.  Here every exchanged value carries its own sequence number: a row publishes m = M^-1 w as ONE
//      16-byte word (mx, my, mz, seq) and a CTA its partial dot products as (gamma, delta, -, seq).  The rows a CTA's
//      products read lie in a few 16-row segments of the vector (nodes that share surface points are neighbours); at its
//      start the kernel lists the segments each CTA needs, and per iteration the CTA's threads fetch exactly those segments
//      and the per-CTA partial sums -- contiguous 256-byte pieces, one word per thread -- re-loading a word until it carries
//      this iteration's sequence number, into shared memory; the row products then run from shared memory.  Data and tag
//      travel in the same naturally aligned 16-byte access, so no fence, no acquire / release pair and no counter is on the
//      path: one L2 round trip per iteration.  (Letting every lane poll the words of its own matrix entries instead --
//      the first version -- flooded L2 with 32 KB of scattered reads per SM per poll and was slower than the barrier.)
//      Two buffers alternate; a CTA can be at most one exchange ahead of the slowest one, because its next publication
//      needs everybody's partial sums of the current step (so nobody still reads the buffer it overwrites).  Sequence
//      numbers are handed out by the host and never repeat, so nothing is cleared between launches.
//      (A thread-block cluster + DSMEM version was measured first -- scratch/cluster_bench.cu, profiles/r02_cluster_bench.txt:
//      barrier.cluster 0.5 k cycles, but the row products of 4096 x 26 non-zeros from the shared memory of 16 SMs cost
//      2-2.6 k cycles per iteration and every pushed halo float4 another 350 per 8 KB; the whole grid through L2 is faster.)
//
//  (2) Assembly balanced inside the CTA and with a quarter of the shared-memory atomics.  3r gave each warp its two rows;
//      the CTA then waited for its slowest warp as long again as the average warp worked.  Here the CTA's (at most 32) rows
//      share one set of shared-memory accumulators and the (row, 64-entry chunk) work units are handed out by a ticket, so
//      all 16 warps finish together.  A contribution is ONE native 32-bit integer atomic (fixed point, scaled per row by
//      the number of contributions it can receive -- integer adds are associative, hence bit-reproducible whatever the
//      arrival order) instead of two; the diagonal (hit by every entry of the list: a 32-way conflict) and the right-hand
//      side are accumulated in registers in 2^40 fixed point and reduced once per chunk.  (Accumulating only the upper triangle of the
//      symmetric matrix and fetching the twins behind the next barrier halves the atomics again -- measured: no gain, the
//      phase is bound by the L2 traffic of its gathers (88 B per list entry), not by the atomics.)
//
// Everything else (residuals / Tukey, regularisation, convergence logic, the arithmetic of the iteration) is 3r's, so both
// produce the same iterates up to the rounding of the matrix entries (2^-24 .. 2^-21 absolute instead of 2^-40).
constexpr int P4_ACC_CAP = 6144;        // 32-bit accumulators per CTA (24 KB)
constexpr int P4_CHUNK = 64;            // transposed-list entries per work unit (2 per lane in flight)
constexpr int P4_ROWS = 2 * (PTPB / 32);  // rows per CTA (P3_R per warp)
constexpr int P4_MAX_N = P3_R * 148 * (PTPB / 32);  // rows the register slots of a full B200 grid hold (4736)
constexpr int P4_MAX_SEG = (P4_MAX_N + 15) / 16;    // 16-row segments of the exchanged vector
constexpr int P4_MAX_CTAS = 256;                   // per-CTA partial sums staged in shared memory
constexpr size_t P4_PQ_BYTES = 2 * (size_t) P4_MAX_CTAS * P4_MAX_CTAS * sizeof(float4);
constexpr unsigned P4_SPIN_LIMIT = 1u << 20;  // re-loads before a missing word is declared lost (seconds: a hang otherwise)

DFU_DEV float4 ld_tagged(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
DFU_DEV void st_tagged(float4* p, float x, float y, float z, unsigned tag) {
    asm volatile("st.relaxed.gpu.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(__uint_as_float(tag)) : "memory");
}
DFU_DEV void warp_sum_ll4(long long& a, long long& b, long long& c, long long& d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
        d += __shfl_xor_sync(0xffffffffu, d, o);
    }
}

// Sums of six per-lane values (two rows x three coordinates) over the warp with 9 shuffles instead of 30: at each of the
// first three butterfly steps a lane keeps one half of its values and hands the other half to its partner, so the values
// spread over the lanes while the sums narrow; the total of value k ends in lanes 4k .. 4k+3.  Returns, in lanes c < 3,
// out0 = sum of v[c] and out1 = sum of v[3 + c].  (Fixed association: deterministic.)
DFU_DEV void warp_reduce6(const float (&v)[6], int lane, float& out0, float& out1) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float lo = v[i], hi = i + 4 < 6 ? v[i + 4] : 0.f;
        const float keep = b4 ? hi : lo, give = b4 ? lo : hi;
        w[i] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
    }
    float x[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = b3 ? w[i + 2] : w[i], give = b3 ? w[i] : w[i + 2];
        x[i] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
    }
    float y = (b2 ? x[1] : x[0]) + __shfl_xor_sync(0xffffffffu, b2 ? x[0] : x[1], 4);
    y += __shfl_xor_sync(0xffffffffu, y, 2);
    y += __shfl_xor_sync(0xffffffffu, y, 1);
    const int c = lane < 3 ? lane : 0;
    out0 = __shfl_sync(0xffffffffu, y, 4 * c);
    out1 = __shfl_sync(0xffffffffu, y, 4 * (3 + c));
}

// residuals + Tukey, t as float4 per node.  P is rarely a multiple of the thread count (75 852 points, 75 776 threads): a
// second dependent round for a handful of points would double the phase.  When the remainder is small it is served by 8
// lanes per point (one neighbour each) whose loads travel together with the last full round's; the 8 products are then
// brought to the group's first lane and summed there in the same order as everywhere else (bit-identical results).
DFU_DEV double phase_point_residual4(const Problem& pb, const float4* __restrict__ t4, bool update_tukey, int tid, int nthreads) {
    double e2 = 0.0;
    const int P = pb.P;
    const int full = P / nthreads, rem = P - full * nthreads;
    const bool rem_split = rem > 0 && full > 0 && rem * 8 <= nthreads;
    const int rounds = full + ((rem > 0 && !rem_split) ? 1 : 0);
    auto finish = [&](int v, float sx, float sy, float sz, float d0, float d1, float d2, float th_old) {
        const float ex = d0 - sx, ey = d1 - sy, ez = d2 - sz;
        float th;
        if (update_tukey) {
            th = tukey_biweight(pb.tukey_offset, pb.psi_data, ex, ey, ez);
            pb.theta[v] = th;
        } else {
            th = th_old;
        }
        pb.s4[v] = make_float4(th * ex, th * ey, th * ez, th);
        e2 += (double) th * ((double) ex * ex + (double) ey * ey + (double) ez * ez);
    };
    for (int round = 0; round < rounds; ++round) {
        const int v = round * nthreads + tid;
        const bool ok = v < P;
        const bool extra_round = rem_split && round == full - 1;  // uniform
        const int pe = full * nthreads + (tid >> 3), ke = tid & 7;
        const bool eok = extra_round && (tid >> 3) < rem;
        int nb[8];
        float w[8];
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, th_old = 0.f;
        int enb = 0;
        float ew = 0.f;
        if (ok) {
            load8(pb.nbr, pb.wts, v, nb, w);
            d0 = pb.dvec[3 * (size_t) v]; d1 = pb.dvec[3 * (size_t) v + 1]; d2 = pb.dvec[3 * (size_t) v + 2];
            if (!update_tukey) th_old = pb.theta[v];
        }
        float ed0 = 0.f, ed1 = 0.f, ed2 = 0.f, eth = 0.f;
        if (eok) {
            enb = pb.nbr[(size_t) pe * 8 + ke];
            ew = pb.wts[(size_t) pe * 8 + ke];
            if (ke == 0) {
                ed0 = pb.dvec[3 * (size_t) pe]; ed1 = pb.dvec[3 * (size_t) pe + 1]; ed2 = pb.dvec[3 * (size_t) pe + 2];
                if (!update_tukey) eth = pb.theta[pe];
            }
        }
        float4 tk[8];
        float4 et = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) {
#pragma unroll
            for (int k = 0; k < 8; ++k) tk[k] = t4[nb[k]];  // (read after an acquiring grid barrier: L1 is clean)
        }
        if (eok) et = t4[enb];
        if (ok) {
            float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                sx = __fmaf_rn(w[k], tk[k].x, sx);
                sy = __fmaf_rn(w[k], tk[k].y, sy);
                sz = __fmaf_rn(w[k], tk[k].z, sz);
            }
            finish(v, sx, sy, sz, d0, d1, d2, th_old);
        }
        if (extra_round) {  // whole warps take this branch
            float sx = 0.f, sy = 0.f, sz = 0.f;
            const int g0 = (threadIdx.x & 31) & ~7;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float wk = __shfl_sync(0xffffffffu, ew, g0 + k);
                const float tx = __shfl_sync(0xffffffffu, et.x, g0 + k), ty = __shfl_sync(0xffffffffu, et.y, g0 + k),
                            tz = __shfl_sync(0xffffffffu, et.z, g0 + k);
                sx = __fmaf_rn(wk, tx, sx);
                sy = __fmaf_rn(wk, ty, sy);
                sz = __fmaf_rn(wk, tz, sz);
            }
            if (eok && ke == 0) finish(pe, sx, sy, sz, ed0, ed1, ed2, eth);
        }
    }
    return e2;
}

// regularisation gather on the float4 copy of the unknowns (same sums as node_gather_reg)
DFU_DEV void node_gather_reg4(const Problem& pb, int n, int lane, const float4* __restrict__ x, float& gx, float& gy, float& gz,
                              float& cnt, float& e2) {
    gx = gy = gz = cnt = e2 = 0.f;
    const float4 xn = x[n];
    const int lo = pb.rin_ptr[n], hi = pb.rin_ptr[n + 1];
    for (int j = lane; j < 8 + (hi - lo); j += 32) {
        const bool out = j < 8;
        const int m = out ? pb.nnbr[(size_t) n * 8 + j] : pb.rin[lo + j - 8];
        if (m == n) continue;
        const float4 xm = x[m];
        const float d0 = xn.x - xm.x, d1 = xn.y - xm.y, d2 = xn.z - xm.z;
        gx += d0; gy += d1; gz += d2;
        cnt += 1.f;
        if (out) e2 += d0 * d0 + d1 * d1 + d2 * d2;
    }
}

struct Exchange4 {
    float4* xt;     // [2][N]            (mx, my, mz, seq)
    float4* pq;     // [2][P4_MAX_CTAS consumer][P4_MAX_CTAS producer] (gamma, delta, -, seq): every CTA has its own inbox
    float4* t4;     // [N]               the unknowns as float4 (x, y, z, 0)
    unsigned seq0;  // first sequence number of this launch (never 0, never re-used)
};

__global__ void __launch_bounds__(PTPB, 1) k_solve_persistent4(Problem pb, Pattern pt, Exchange4 ex4, SolveCtl ctl, Scalars* sc,
                                                               unsigned* bar) {
    constexpr int NWARP = PTPB / 32;
    __shared__ double shw[3 * NWARP];
    __shared__ double tot_sm[3];
    __shared__ __align__(8) unsigned s_acc[P4_ACC_CAP];
    __shared__ unsigned long long s_rsum[P4_ROWS][4];  // per local row: b.x, b.y, b.z, diagonal (2^40 fixed point)
    __shared__ int s_lo[P4_ROWS], s_deg[P4_ROWS], s_ds[P4_ROWS], s_len[P4_ROWS];
    __shared__ int s_rowbase[P4_ROWS + 1], s_unitbase[P4_ROWS + 1];
    __shared__ int s_ticket, s_fail, s_nseg;
    __shared__ unsigned s_need[(P4_MAX_N + 31) / 32];      // nodes whose exchanged word this CTA's row products read
    __shared__ unsigned short s_segpos[P4_MAX_SEG];        // 16-row segment -> its position in xs
    __shared__ unsigned short s_seglist[P4_MAX_SEG];       // the segments this CTA fetches
    __shared__ float2 s_ps[P4_MAX_CTAS];                   // the per-CTA partial sums of the current exchange
    extern __shared__ float4 xs[];                         // fetched segments: [s_nseg][16] (mx, my, mz, seq)
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, gw = tid >> 5, nw = nthreads >> 5;
    const int nb = gridDim.x, N = pb.N;
    unsigned bar_target = 0;
    unsigned seq = ex4.seq0;
#define GRID_SYNC() grid_barrier(bar, (unsigned) nb, bar_target)
#define PART(buf) (pb.part + (size_t) (buf) * 2 * MAX_PARTIALS)
    // per-phase SM cycles of thread 0 (DFU_SOLVER_PROFILE): accumulated in shared memory, copied out at the end (a global
    // read-modify-write per sample would charge an L2 round trip to the next phase)
    __shared__ long long s_prof[16];
    if (ctl.prof && threadIdx.x < 16) s_prof[threadIdx.x] = 0;
    long long t_prev = clock64();
#define PROF(k)                                               \
    do {                                                      \
        if (ctl.prof && tid == 0) {                           \
            const long long t_now = clock64();                \
            s_prof[k] += t_now - t_prev;                      \
            t_prev = t_now;                                   \
        }                                                     \
    } while (0)
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
    auto publish = [&](double a, double b, double c, double* dst) {
        a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
        __syncthreads();
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

    // ---- my rows (warp gw owns rows gw and gw + nw; local row index lr = r * NWARP + wib) ------------------------
    int rn[P3_R], roff[P3_R], rlen[P3_R], rds[P3_R];
    int rc[P3_R][P3_LE];
    float ra[P3_R][P3_LE], rv[P3_R][P3_LE];
    float rinv[P3_R];
    double rinvd[P3_R];
    float s_r[P3_R], s_w[P3_R], s_z[P3_R], s_s[P3_R], s_p[P3_R], s_x[P3_R], s_t[P3_R];  // coordinate `lane` (lanes 0..2)
#pragma unroll
    for (int r = 0; r < P3_R; ++r) {
        const int n = gw + r * nw;
        const int lr = r * NWARP + wib;
        rn[r] = n < N ? n : -1;
        roff[r] = rlen[r] = rds[r] = 0;
        rinv[r] = 0.f;
        rinvd[r] = 0.0;
        s_r[r] = s_w[r] = s_z[r] = s_s[r] = s_p[r] = s_x[r] = s_t[r] = 0.f;
#pragma unroll
        for (int u = 0; u < P3_LE; ++u) {
            rc[r][u] = -1;
            ra[r][u] = rv[r][u] = 0.f;
        }
        int lo = 0, deg = 0;
        if (n < N) {
            roff[r] = pt.rowptr[n];
            rlen[r] = pt.rowlen[n];
            rds[r] = pt.dslot[n];
            lo = pb.tptr[n];
            deg = pb.tptr[n + 1] - lo;
#pragma unroll
            for (int u = 0; u < P3_LE; ++u) {
                const int j = lane + 32 * u;
                if (j < rlen[r]) {
                    rc[r][u] = pt.col[roff[r] + j];
                    ra[r][u] = pt.areg[roff[r] + j];
                }
            }
            if (lane < 3) pb.t[3 * (size_t) n + lane] = 0.f;  // unknowns := 0 (opt_solver.cpp:192-193)
            if (lane == 0) ex4.t4[n] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (lane == 0) {
            s_lo[lr] = lo; s_deg[lr] = deg; s_ds[lr] = rds[r];
            // rows that can collect 2048 or more contributions per column use 64-bit accumulators (two slots per column)
            s_len[lr] = n < N ? (deg > FIX_MAX_DEG ? 2 * rlen[r] : rlen[r]) : 0;
        }
    }
    if (threadIdx.x == 0) {
        s_fail = 0;
        s_nseg = 0;
    }
    if (tid == 0) sc->spin_fail = 0;
    for (int i = threadIdx.x; i < (P4_MAX_N + 31) / 32; i += blockDim.x) s_need[i] = 0u;
    __syncthreads();
    // which 16-row segments of the exchanged vector do this CTA's row products read?
#pragma unroll
    for (int r = 0; r < P3_R; ++r) {
        if (rn[r] < 0) continue;
#pragma unroll
        for (int u = 0; u < P3_LE; ++u)
            if (rc[r][u] >= 0) atomicOr(&s_need[rc[r][u] >> 5], 1u << (rc[r][u] & 31));
        for (int j = 32 * P3_LE + lane; j < rlen[r]; j += 32) {
            const int c = pt.col[roff[r] + j];
            atomicOr(&s_need[c >> 5], 1u << (c & 31));
        }
    }
    __syncthreads();
    for (int sg = threadIdx.x; sg < (N + 15) / 16; sg += blockDim.x)
        if ((s_need[sg >> 1] >> ((sg & 1) * 16)) & 0xffffu) {
            const int pos = atomicAdd(&s_nseg, 1);  // (any order: only the mapping depends on it)
            s_segpos[sg] = (unsigned short) pos;
            s_seglist[pos] = (unsigned short) sg;
        }
    __syncthreads();
    // columns -> positions in xs
#pragma unroll
    for (int r = 0; r < P3_R; ++r)
#pragma unroll
        for (int u = 0; u < P3_LE; ++u)
            if (rc[r][u] >= 0) rc[r][u] = (int) s_segpos[rc[r][u] >> 4] * 16 + (rc[r][u] & 15);
    // accumulator layout: all rows of the CTA at once when they fit, one row per pass otherwise (a row of N <= 4736 columns
    // always fits); work units = 64-entry chunks of the rows' transposed lists
    bool one_pass;
    {
        int tot = 0;
        for (int i = 0; i < P4_ROWS; ++i) tot += s_len[i];
        one_pass = tot + P4_ROWS <= P4_ACC_CAP;  // (+ one slot of alignment padding per row)
    }
    GRID_SYNC();

    // Exchange step: every thread fetches its words of the segments this CTA needs (and thread c < nb the partial sums of
    // CTA c) until they carry the sequence number `tag`, into shared memory.  Ends with a CTA barrier.
    auto fetch = [&](const float4* xg, const float4* pg, unsigned tag, bool want_x) {
        const int nent = want_x ? 16 * s_nseg : 0;
        unsigned spins = 0;
        const bool has_p = pg != nullptr && (int) threadIdx.x < nb;
        float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_p) pv = ld_tagged(pg + threadIdx.x);
        for (int e0 = threadIdx.x; e0 < nent; e0 += 4 * PTPB) {
            const float4* src[4];
            float4 v[4];
            bool pend[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int e = e0 + q * PTPB;
                const int node = e < nent ? (int) s_seglist[e >> 4] * 16 + (e & 15) : N;
                pend[q] = node < N;  // (the last segment may reach beyond the last row)
                src[q] = xg + (pend[q] ? node : 0);
                v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pend[q]) v[q] = ld_tagged(src[q]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (pend[q]) {
                    while (__float_as_uint(v[q].w) != tag && ++spins < P4_SPIN_LIMIT) v[q] = ld_tagged(src[q]);
                    if (__float_as_uint(v[q].w) != tag) s_fail = 1;
                }
                if (e0 + q * PTPB < nent) xs[e0 + q * PTPB] = v[q];
            }
        }
        if (has_p) {
            while (__float_as_uint(pv.w) != tag && ++spins < P4_SPIN_LIMIT) pv = ld_tagged(pg + threadIdx.x);
            if (__float_as_uint(pv.w) != tag) s_fail = 1;
            s_ps[threadIdx.x] = make_float2(pv.x, pv.y);
        }
        __syncthreads();
    };
    // row products with the fetched vector (shared memory); entries beyond the register slots take their column from L2
    auto spmv_xs = [&](float (&out)[P3_R]) {
        static_assert(P3_R == 2, "warp_reduce6 sums two rows");
        float acc[6];
#pragma unroll
        for (int r = 0; r < P3_R; ++r) {
            float ax = 0.f, ay = 0.f, az = 0.f;
            if (rn[r] >= 0) {
#pragma unroll
                for (int u = 0; u < P3_LE; ++u)
                    if (rc[r][u] >= 0) {
                        const float4 m = xs[rc[r][u]];
                        ax = __fmaf_rn(rv[r][u], m.x, ax);
                        ay = __fmaf_rn(rv[r][u], m.y, ay);
                        az = __fmaf_rn(rv[r][u], m.z, az);
                    }
                for (int j = 32 * P3_LE + lane; j < rlen[r]; j += 32) {
                    const float v = pt.vals[roff[r] + j];
                    const int c = pt.col[roff[r] + j];
                    const float4 m = xs[(int) s_segpos[c >> 4] * 16 + (c & 15)];
                    ax = __fmaf_rn(v, m.x, ax);
                    ay = __fmaf_rn(v, m.y, ay);
                    az = __fmaf_rn(v, m.z, az);
                }
            }
            acc[3 * r] = ax; acc[3 * r + 1] = ay; acc[3 * r + 2] = az;
        }
        warp_reduce6(acc, lane, out[0], out[1]);
    };
    // lane 0 of a warp publishes the warp's rows of a vector held by lanes 0..2 (one coordinate each)
    auto publish_rows = [&](float4* xg, const float (&val)[P3_R], unsigned tag) {
#pragma unroll
        for (int r = 0; r < P3_R; ++r) {
            if (rn[r] < 0) continue;
            const float vy = __shfl_sync(0xffffffffu, val[r], 1), vz = __shfl_sync(0xffffffffu, val[r], 2);
            if (lane == 0) st_tagged(xg + rn[r], val[r], vy, vz, tag);
        }
    };

    double rz_ref = -1.0, E = 0.0, E0 = 0.0;
    int pcg_total = 0, gn_total = 0;
    bool first = true, stop_all = false;

    for (int outer = 0; outer < ctl.num_iter && !stop_all; ++outer) {
        for (int gn = 0; gn < ctl.nonlinear_iter; ++gn) {
            PROF(0);
            const double e2_local = phase_point_residual4(pb, ex4.t4, gn == 0, tid, nthreads);
            PROF(1);
            GRID_SYNC();
            PROF(2);
            // ---- b = sum tw * theta e and (when theta changed) the data part of the CTA's rows of A ----------------
            const bool assemble = gn == 0;
            const int npass = one_pass ? 1 : P4_ROWS;
            for (int pass = 0; pass < npass; ++pass) {
                const int lr0 = one_pass ? 0 : pass, lr1 = one_pass ? P4_ROWS : pass + 1;
                if (wib == 0) {  // accumulator offsets and work-unit prefix of the rows of this pass
                    const int lr = lr0 + lane;
                    const bool in = lr < lr1;
                    int len = in ? s_len[lr] : 0, units = in ? (s_deg[lr] + P4_CHUNK - 1) / P4_CHUNK : 0;
                    len += len & 1;  // keep 64-bit views aligned
                    int il = len, iu = units;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int a = __shfl_up_sync(0xffffffffu, il, o), b2 = __shfl_up_sync(0xffffffffu, iu, o);
                        if (lane >= o) {
                            il += a;
                            iu += b2;
                        }
                    }
                    if (in) {
                        s_rowbase[lr] = il - len;
                        s_unitbase[lr] = iu - units;
                    }
                    if (lane == 31) {
                        s_rowbase[P4_ROWS] = il;
                        s_unitbase[P4_ROWS] = iu;
                        s_ticket = 0;
                    }
                }
                for (int i = threadIdx.x; i < 4 * P4_ROWS; i += blockDim.x) (&s_rsum[0][0])[i] = 0ull;
                __syncthreads();
                if (assemble)
                    for (int i = threadIdx.x; i < s_rowbase[P4_ROWS]; i += blockDim.x) s_acc[i] = 0u;
                __syncthreads();
                const int n_units = s_unitbase[P4_ROWS];
                for (;;) {
                    int u = 0;
                    if (lane == 0) u = atomicAdd(&s_ticket, 1);
                    u = __shfl_sync(0xffffffffu, u, 0);
                    if (u >= n_units) break;
                    // unit -> (local row, chunk): the last row whose first unit is <= u
                    const int lrq = lr0 + lane;
                    const unsigned le = __ballot_sync(0xffffffffu, lrq < lr1 && s_unitbase[lrq] <= u && s_deg[lrq] > 0);
                    const int lr = lr0 + 31 - __clz(le);
                    const int chunk = u - s_unitbase[lr];
                    const int lo = s_lo[lr], deg = s_deg[lr], ds = s_ds[lr], base = s_rowbase[lr];
                    const bool wide = deg > FIX_MAX_DEG;
                    // per-row fixed-point scale of the 32-bit accumulators: deg < 2^bits contributions <= 1 each
                    const int bits = 32 - __clz(deg);
                    const float scale = __uint_as_float((unsigned) (127 + 32 - bits) << 23);
                    long long bx = 0, by = 0, bz = 0, dg = 0;
                    int v[2];
                    float w[2];
                    uint4 sl[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int e = chunk * P4_CHUNK + lane + 32 * q;
                        const bool ok = e < deg;
                        v[q] = ok ? pb.tv[lo + e] : -1;
                        w[q] = ok ? pb.tw[lo + e] : 0.f;
                        sl[q] = (ok && assemble) ? pt.tslot[lo + e] : make_uint4(0u, 0u, 0u, 0u);
                    }
                    float th[2];
                    float4 w0[2], w1[2], s4v[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int vv = v[q] >= 0 ? v[q] : 0;
                        s4v[q] = pb.s4[vv];
                        th[q] = s4v[q].w;  // s4 = (theta e, theta)
                        if (assemble) {
                            w0[q] = *(reinterpret_cast<const float4*>(pb.wts) + 2 * (size_t) vv);
                            w1[q] = *(reinterpret_cast<const float4*>(pb.wts) + 2 * (size_t) vv + 1);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        if (v[q] < 0) continue;
                        bx += __float2ll_rn(w[q] * s4v[q].x * FIX_SCALE);
                        by += __float2ll_rn(w[q] * s4v[q].y * FIX_SCALE);
                        bz += __float2ll_rn(w[q] * s4v[q].z * FIX_SCALE);
                        const float c = th[q] * w[q];
                        if (!assemble || c == 0.f) continue;
                        dg += __float2ll_rn(c * w[q] * FIX_SCALE);
                        const float wk[8] = {w0[q].x, w0[q].y, w0[q].z, w0[q].w, w1[q].x, w1[q].y, w1[q].z, w1[q].w};
                        const unsigned sk[8] = {sl[q].x & 0xffffu, sl[q].x >> 16, sl[q].y & 0xffffu, sl[q].y >> 16,
                                                sl[q].z & 0xffffu, sl[q].z >> 16, sl[q].w & 0xffffu, sl[q].w >> 16};
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            if ((int) sk[k] == ds) continue;  // the diagonal is summed in registers (dg)
                            if (wide)
                                atomicAdd(reinterpret_cast<unsigned long long*>(s_acc + base) + sk[k],
                                          (unsigned long long) __float2ll_rn(c * wk[k] * FIX_SCALE));
                            else
                                atomicAdd(&s_acc[base + sk[k]], __float2uint_rn(c * wk[k] * scale));
                        }
                    }
                    warp_sum_ll4(bx, by, bz, dg);
                    if (lane == 0) {
                        atomicAdd(&s_rsum[lr][0], (unsigned long long) bx);
                        atomicAdd(&s_rsum[lr][1], (unsigned long long) by);
                        atomicAdd(&s_rsum[lr][2], (unsigned long long) bz);
                        atomicAdd(&s_rsum[lr][3], (unsigned long long) dg);
                    }
                }
                __syncthreads();
                PROF(3);
                // ---- every warp collects its own rows of this pass -------------------------------------------------
#pragma unroll
                for (int r = 0; r < P3_R; ++r) {
                    const int lr = r * NWARP + wib;
                    if (rn[r] < 0 || lr < lr0 || lr >= lr1) continue;  // uniform over the warp
                    if (assemble) {
                        const int base = s_rowbase[lr], deg = s_deg[lr];
                        const bool wide = deg > FIX_MAX_DEG;
                        const int bits = 32 - __clz(deg);
                        const double inv_scale = 1.0 / (double) __uint_as_float((unsigned) (127 + 32 - bits) << 23);
                        const float dgf = (float) ((double) (long long) s_rsum[lr][3] * FIX_INV);
                        auto value = [&](int j) -> float {
                            if (j == rds[r]) return dgf;
                            if (deg == 0) return 0.f;
                            return wide ? (float) ((double) (long long) (reinterpret_cast<unsigned long long*>(s_acc + base))[j] * FIX_INV)
                                        : (float) ((double) s_acc[base + j] * inv_scale);
                        };
#pragma unroll
                        for (int u = 0; u < P3_LE; ++u)
                            if (rc[r][u] >= 0) rv[r][u] = ra[r][u] + value(lane + 32 * u);
                        for (int j = 32 * P3_LE + lane; j < rlen[r]; j += 32) pt.vals[roff[r] + j] = pt.areg[roff[r] + j] + value(j);
                        float D;
                        if (rds[r] < 32 * P3_LE) {
                            float pick = rv[r][0];
#pragma unroll
                            for (int u = 1; u < P3_LE; ++u) pick = (rds[r] >> 5) == u ? rv[r][u] : pick;
                            D = __shfl_sync(0xffffffffu, pick, rds[r] & 31);
                        } else {
                            D = pt.areg[roff[r] + rds[r]] + dgf;
                        }
                        rinv[r] = D > 0.f ? 1.f / D : 0.f;
                        rinvd[r] = D > 0.f ? 1.0 / (double) D : 0.0;
                        if (lane == 0) pb.nbuf[3 * (size_t) N + rn[r]] = D;
                    }
                    // right-hand side of the row: lanes 0..2 take their coordinate
                    const float bdat = lane < 3 ? (float) ((double) (long long) s_rsum[lr][lane] * FIX_INV) : 0.f;
                    s_r[r] = bdat;
                }
                if (!one_pass) __syncthreads();  // the accumulators are re-used by the next pass
            }
            double rz = 0.0, er = 0.0;
#pragma unroll
            for (int r = 0; r < P3_R; ++r) {
                if (rn[r] < 0) continue;
                const int n = rn[r];
                float b = s_r[r];
                if (pb.wreg2 > 0.f) {
                    float gx, gy, gz, cnt, e2;
                    node_gather_reg4(pb, n, lane, ex4.t4, gx, gy, gz, cnt, e2);
                    gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
                    e2 = warp_sum(e2);
                    const float g = lane == 0 ? gx : (lane == 1 ? gy : gz);
                    b -= pb.wreg2 * g;
                    if (lane == 0) er += (double) pb.wreg2 * e2;
                }
                if (lane < 3) {
                    pb.nbuf[3 * (size_t) n + lane] = b;
                    s_r[r] = b;
                    s_x[r] = 0.f;
                    rz += (double) b * (double) b * rinvd[r];
                } else {
                    s_r[r] = 0.f;
                }
            }
            ++seq;  // u0 = M^-1 b travels like the iterates
            {
                float u0[P3_R];
#pragma unroll
                for (int r = 0; r < P3_R; ++r) u0[r] = s_r[r] * rinv[r];
                publish_rows(ex4.xt + (size_t) (seq & 1u) * N, u0, seq);
            }
            PROF(4);
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
                {  // w0 = A u0 (the barrier above has made every row's u0 visible; the tags only confirm it)
                    fetch(ex4.xt + (size_t) (seq & 1u) * N, nullptr, seq, true);
                    float w0[P3_R];
                    spmv_xs(w0);
#pragma unroll
                    for (int r = 0; r < P3_R; ++r) {
                        s_w[r] = (rn[r] >= 0 && lane < 3) ? w0[r] : 0.f;
                        s_z[r] = s_s[r] = s_p[r] = 0.f;
                    }
                }
                float inv_gamma_prev = 0.f, inv_alpha_prev = 0.f;
                PROF(8);
                for (int it = 0; it < ctl.linear_iter; ++it) {
                    ++seq;
                    const int buf = (int) (seq & 1u);
                    float4* xt = ex4.xt + (size_t) buf * N;
                    float4* pq = ex4.pq + (size_t) buf * P4_MAX_CTAS * P4_MAX_CTAS;
                    const bool last = it + 1 >= ctl.linear_iter;
                    // (r,u), (w,u) with u = M^-1 r; m = M^-1 w is published
                    float g = 0.f, d = 0.f;
                    float mv[P3_R];
#pragma unroll
                    for (int r = 0; r < P3_R; ++r) {
                        mv[r] = 0.f;
                        if (rn[r] < 0 || lane >= 3) continue;
                        const float ur = s_r[r] * rinv[r];
                        g = __fmaf_rn(s_r[r], ur, g);
                        d = __fmaf_rn(s_w[r], ur, d);
                        mv[r] = s_w[r] * rinv[r];
                    }
                    if (!last) publish_rows(xt, mv, seq);
                    g += __shfl_xor_sync(0xffffffffu, g, 1); d += __shfl_xor_sync(0xffffffffu, d, 1);
                    g += __shfl_xor_sync(0xffffffffu, g, 2); d += __shfl_xor_sync(0xffffffffu, d, 2);
                    if (lane == 0) reinterpret_cast<float2*>(shw)[wib] = make_float2(g, d);
                    __syncthreads();
                    PROF(9);
                    if (wib == 0) {  // CTA total in a fixed order (xor tree over the 16 warps), published with its tag
                        float2 t = lane < NWARP ? reinterpret_cast<float2*>(shw)[lane] : make_float2(0.f, 0.f);
#pragma unroll
                        for (int o = NWARP / 2; o > 0; o >>= 1) {
                            t.x += __shfl_xor_sync(0xffffffffu, t.x, o);
                            t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
                        }
                        // pushed into every CTA's own inbox: nobody polls a line that somebody else polls too (148 CTAs reading
                        // the same 19 lines of a shared array serialise in their L2 slices: 2.8 k cycles in 3r, worse when polled)
                        t.x = __shfl_sync(0xffffffffu, t.x, 0);
                        t.y = __shfl_sync(0xffffffffu, t.y, 0);
                        for (int c = lane; c < nb; c += 32) st_tagged(pq + (size_t) c * P4_MAX_CTAS + blockIdx.x, t.x, t.y, 0.f, seq);
                    }
                    // one round trip: the segments of m this CTA reads and everybody's partial sums
                    fetch(xt, pq + (size_t) blockIdx.x * P4_MAX_CTAS, seq, !last);
                    PROF(10);
                    // every warp sums the partials in the same fixed order (no broadcast needed), then its row products
                    float gs = 0.f, ds = 0.f;
                    for (int i = lane; i < nb; i += 32) {
                        const float2 x = s_ps[i];
                        gs += x.x; ds += x.y;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        gs += __shfl_xor_sync(0xffffffffu, gs, o);
                        ds += __shfl_xor_sync(0xffffffffu, ds, o);
                    }
                    const double gamma = (double) gs, delta = (double) ds;
                    float nv[P3_R];
#pragma unroll
                    for (int r = 0; r < P3_R; ++r) nv[r] = 0.f;
                    if (!last) spmv_xs(nv);
                    PROF(11);
                    if (!(gamma > 0.0) || (it > 0 && gamma <= ctl.tol2 * rz_ref)) break;
                    // beta = gamma / gamma_prev, alpha = gamma / (delta - beta gamma / alpha_prev).  The partial sums are floats
                    // and alpha, beta are applied as floats, so the quotients are float reciprocals (two dependent double
                    // divisions cost more than the row products); the difference, which cancels, is formed in double from the
                    // values actually applied
                    const float bf = it > 0 ? gs * inv_gamma_prev : 0.f;
                    const double denom = it > 0 ? delta - (double) bf * gamma * (double) inv_alpha_prev : delta;
                    if (!(denom > 0.0)) break;
                    const float af = gs * __frcp_rn((float) denom);
                    if (!(af > 0.f)) break;
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
                    inv_gamma_prev = __frcp_rn(gs);
                    inv_alpha_prev = __frcp_rn(af);
                }
            }
#pragma unroll
            for (int r = 0; r < P3_R; ++r) {
                if (rn[r] < 0) continue;
                if (lane < 3) {
                    s_t[r] += s_x[r];
                    pb.t[3 * (size_t) rn[r] + lane] = s_t[r];
                }
                const float ty = __shfl_sync(0xffffffffu, s_t[r], 1), tz = __shfl_sync(0xffffffffu, s_t[r], 2);
                if (lane == 0) ex4.t4[rn[r]] = make_float4(s_t[r], ty, tz, 0.f);
            }
            ++gn_total;
            PROF(13);
            GRID_SYNC();
            PROF(14);
        }
    }
    // ---- final energy at the solution, Tukey weights of the last outer iteration -------------------------
    {
        const double e2 = phase_point_residual4(pb, ex4.t4, first, tid, nthreads);  // no GN step ran: weights at t = 0
        double er = 0.0;
        if (pb.wreg2 > 0.f) {
#pragma unroll
            for (int r = 0; r < P3_R; ++r) {
                if (rn[r] < 0) continue;
                float gx, gy, gz, cnt, r2;
                node_gather_reg4(pb, rn[r], lane, ex4.t4, gx, gy, gz, cnt, r2);
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
    if (threadIdx.x == 0 && s_fail) sc->spin_fail = 1;
    if (tid == 0) {
        sc->E = E;
        sc->E0 = first ? E : E0;
        sc->rz_ref = rz_ref;
        sc->pcg_iters = pcg_total;
        sc->gn_steps = gn_total;
        sc->first = 0;
    }
    PROF(15);
    if (ctl.prof && tid == 0)
        for (int i = 0; i < 16; ++i) ctl.prof[i] = s_prof[i];
#undef PROF
#undef PART
#undef GRID_SYNC
}
