// Solver: one kernel per phase (data-parallel ranks with an all-reduce between phases)
// (textually included by solver.cu inside its anonymous namespace -- one translation unit)
#pragma once

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

