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
// Execution paths, same mathematics (tests run every one of them against the oracle):
//   * solver_normal.cuh   version 3 / 3r (default for the frame loop): the N x N normal matrix is assembled explicitly
//                         once per GN step (fixed-point, order-independent) and the iteration is the pipelined PCG of
//                         Ghysels & Vanroose with ONE grid barrier per iteration; 3r keeps the rows in registers;
//   * solver_matfree.cuh  versions 1 / 2: matrix-free textbook PCG in one persistent cooperative kernel (3 barriers per
//                         iteration), used for long tolerance-driven solves;
//   * solver_phases.cuh   data-parallel ranks (all-reduce hook / NCCL communicator set): one kernel per phase, nbuf
//                         all-reduced once per GN step and q (3N floats) once per PCG iteration; everything after an
//                         all-reduce is evaluated in a fixed order, so all ranks compute bit-identical iterates.
// Convergence and early-out are decided on the device: no host round trip inside solveAll on the persistent paths.
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

#include "solver_common.cuh"
#include "solver_phases.cuh"
#include "solver_matfree.cuh"
#include "solver_normal.cuh"
#include "solver_v4.cuh"
#include "solver_graph.cuh"
#include "solver_p2plane.cuh"
#include "solver_p2plane_persistent.cuh"

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
    bool bar_clean = false;      // the last kernel on `bar` left it cleared (version 3r does)
    bool writeback_fused = false;  // the last solve wrote the node transforms back itself (version 3r)
    uint64_t reg_epoch = 0;      // node-position epoch the regularisation graph was built for (0: never)
    bool problem_ready = false;
    int coop_blocks = 0;         // co-resident CTAs for the persistent kernel (0: not available)
    int coop_blocks2 = 0;        // same for version 2 of the kernel
    int coop_blocks3 = 0;        // same for version 3
    int coop_blocks3r = 0;       // same for version 3 with the rows in registers
    int coop_blocks4 = 0;        // same for version 4 (tagged exchange, CTA-balanced assembly)
    float4* t4 = nullptr;        // version 4: the unknowns as float4 per node
    unsigned seq = 0;            // version 4: last sequence number handed out (tags of the exchanged words never repeat)
    int coop_blocks_p2p = 0;     // same for the point-to-plane kernel (P2P_TPB threads, up to 2 CTAs per SM)
    int last_kernel = 0;         // which persistent kernel the last solve used (1 / 2 / 3; 5 = version 4; 4 = point-to-plane; 0: multi-kernel path)
    // explicit normal matrix (version 3): pattern per frame, values per re-weighting
    int *rowptr = nullptr, *rowlen = nullptr, *dslot = nullptr, *pat_cursor = nullptr;
    int32_t* col = nullptr;
    float *areg = nullptr, *vals = nullptr;
    uint4* tslot = nullptr;
    float4 *exch = nullptr, *st = nullptr;
    unsigned long long *xw = nullptr, *pw = nullptr;
    size_t cap_nnz = 0, cap_slots = 0, cap_rows = 0;
    bool pattern_ready = false;
    // north-star extension: point-to-plane SE(3) data term (solver_p2plane.cuh)
    int energy_mode = DFU_ENERGY_REF_TRANSLATION;
    int reg_mode = DFU_REG_QUADRATIC;   // regulariser of the point-to-plane energy
    const float *canon_v = nullptr, *live_v = nullptr, *live_n = nullptr;  // caller-owned, valid until solve_all returns
    float *p2p_pt = nullptr, *p2p_node = nullptr;                          // per-point / per-node scratch
    size_t p2p_cap_pt = 0, p2p_cap_node = 0;
    float* p2p_X = nullptr;                                                // the increments X_n inside p2p_node
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
    cudaFree(s->tslot); cudaFree(s->exch); cudaFree(s->st); cudaFree(s->xw); cudaFree(s->pw); cudaFree(s->t4);
    s->xw = s->pw = nullptr;
    s->t4 = nullptr;
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
    // (version 3r leaves the barrier words cleared itself; anything else ran last: clear them here)
    bool bar_cleared = false;
    auto clear_bar = [&]() -> int {
        if (!bar_cleared) DFU_CUDA_OK(cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned), st));
        bar_cleared = true;
        return DFU_OK;
    };
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
    bool ran_v4 = false, fused = false;
    if (ver == 3 && !(s->pattern_ready && s->coop_blocks3 > 0 && s->N <= 32 * s->coop_blocks3 * (PTPB / 32))) ver = 2;  // (lane-per-row blocks)
    if (ver == 2 && (s->coop_blocks2 == 0 || s->N > P2_NPW * s->coop_blocks2 * (PTPB / 32))) ver = 1;
    if (ver == 3) {
        Pattern pt{s->rowptr, s->rowlen, s->dslot, s->col, s->areg, s->vals, s->tslot, s->exch, s->st, s->xw, s->pw, nullptr, nullptr, nullptr, nullptr, nullptr};
        void* args[] = {&pb, &pt, &ctl, &sc, &bar};
        // rows in registers when every node fits a register slot; DFU_SOLVER_PATH=p3g forces the generic kernel, p3 the
        // barrier-per-iteration register kernel (3r), p4 / default: version 4 (tagged exchange, CTA-balanced assembly)
        const bool reg = s->coop_blocks3r > 0 && s->N <= P3_R * s->coop_blocks3r * (PTPB / 32) && !(force && force[1] == '3' && force[2] == 'g');
        // version 4 (barrier-free PCG exchange + CTA-balanced assembly) measures the same as 3r on B200 (profiles/r02_solver_experiments.md):
        // opt-in with DFU_SOLVER_PATH=p4, covered by the same parity tests
        const bool v4 = reg && s->coop_blocks4 > 0 && s->N <= P3_R * s->coop_blocks4 * (PTPB / 32) && s->N <= P4_MAX_N &&
                        s->coop_blocks4 <= P4_MAX_CTAS && force && force[0] == 'p' && force[1] == '4';
        if (v4) {
            // sequence numbers of the tagged words: one per PCG iteration, never re-used (cleared buffers hold 0)
            // (per Gauss-Newton step: one for u0 and one per PCG iteration)
            const unsigned long long need = (unsigned long long) std::max(ctl.num_iter, 0) * std::max(ctl.nonlinear_iter, 0) *
                                                ((unsigned long long) std::max(ctl.linear_iter, 0) + 1) + 2;
            DFU_REQUIRE(need < 0x40000000ull, DFU_ERR_INVALID, "iteration budget too large for the tagged exchange");
            if ((unsigned long long) s->seq + need >= 0xfffffff0ull) {  // wrap-around (after ~10^7 solves): start over on clean buffers
                DFU_CUDA_OK(cudaMemsetAsync(s->xw, 0, 2 * 3 * (size_t) s->N * sizeof(unsigned long long), st));
                DFU_CUDA_OK(cudaMemsetAsync(s->pw, 0, P4_PQ_BYTES, st));
                s->seq = 0;
            }
            Exchange4 ex4{reinterpret_cast<float4*>(s->xw), reinterpret_cast<float4*>(s->pw), s->t4, s->seq};
            s->seq += (unsigned) need;
            void* args4[] = {&pb, &pt, &ex4, &ctl, &sc, &bar};
            if (clear_bar() != DFU_OK) return DFU_ERR_CUDA;
            const size_t xs_bytes = (size_t) ((s->N + 15) / 16) * 16 * sizeof(float4);  // the fetched segments of the exchanged vector
            DFU_CUDA_OK(cudaLaunchCooperativeKernel((void*) k_solve_persistent4, dim3(s->coop_blocks4), dim3(PTPB), args4, xs_bytes, st));
            ran_v4 = true;
        } else if (reg) {
            // one launch per solve and nothing around it: the kernel clears its own barrier words on the way out and writes the
            // node transforms (and the warp field's flags) back itself
            if (!s->bar_clean && clear_bar() != DFU_OK) return DFU_ERR_CUDA;
            pt.wf_real = s->wf->real; pt.wf_dual = s->wf->dual; pt.wf_pos_w = s->wf->pos_w; pt.wf_flags = s->wf->flags;
            pt.t4 = s->t4;
            DFU_CUDA_OK(cudaLaunchCooperativeKernel((void*) k_solve_persistent3r, dim3(s->coop_blocks3r), dim3(PTPB), args, 0, st));
            fused = true;
        } else {
            if (clear_bar() != DFU_OK) return DFU_ERR_CUDA;
            DFU_CUDA_OK(cudaLaunchCooperativeKernel((void*) k_solve_persistent3, dim3(s->coop_blocks3), dim3(PTPB), args, 0, st));
        }
        if (getenv("DFU_DEBUG")) fprintf(stderr, "[dfu] v3 %s\n", reg ? "rows in registers" : "generic");
    } else {
        void* args[] = {&pb, &ctl, &sc, &bar};
        if (clear_bar() != DFU_OK) return DFU_ERR_CUDA;
        if (ver == 1)
            DFU_CUDA_OK(cudaLaunchCooperativeKernel((void*) k_solve_persistent, dim3(s->coop_blocks), dim3(PTPB), args, 0, st));
        else
            DFU_CUDA_OK(cudaLaunchCooperativeKernel((void*) k_solve_persistent2, dim3(s->coop_blocks2), dim3(PTPB), args, 0, st));
    }
    s->last_kernel = ran_v4 ? 5 : ver;
    s->bar_clean = fused;
    s->writeback_fused = fused;
    ++g_dfu_launches;
    if (profile && ver == 3) {  // debugging aid: synchronises
        long long h[16];
        DFU_CUDA_OK(cudaMemcpyAsync(h, prof_dev, sizeof(h), cudaMemcpyDeviceToHost, st));
        DFU_CUDA_OK(cudaStreamSynchronize(st));
        static const char* names3[16] = {"misc", "residual", "bar", "gather_b", "assemble", "row_init", "bar", "sum_part", "spmv_w0",
                                         "dots+m", "bar", "sum_part", "spmv+update", "t+=x", "bar", "final"};
        static const char* names4[16] = {"misc", "residual", "bar", "assemble", "collect+reg", "publish", "bar+totals", "-", "spmv_w0",
                                         "dots+m", "exchange+spmv", "cta_sync", "update", "t+=x", "bar", "final"};
        const char** names = s->last_kernel == 5 ? names4 : names3;
        fprintf(stderr, "[dfu] v%d cycles of CTA 0:", s->last_kernel == 5 ? 4 : 3);
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
    if (s->energy_mode != DFU_ENERGY_REF_TRANSLATION) return false;
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
        cudaFree(s->rowptr); cudaFree(s->rowlen); cudaFree(s->dslot); cudaFree(s->exch); cudaFree(s->st); cudaFree(s->xw); cudaFree(s->t4);
        s->rowptr = s->rowlen = s->dslot = nullptr; s->exch = s->st = nullptr; s->xw = nullptr; s->t4 = nullptr; s->cap_rows = 0;
        DFU_CUDA_OK(cudaMalloc(&s->rowptr, (size_t) N * sizeof(int)));
        DFU_CUDA_OK(cudaMalloc(&s->rowlen, (size_t) N * sizeof(int)));
        DFU_CUDA_OK(cudaMalloc(&s->dslot, (size_t) N * sizeof(int)));
        DFU_CUDA_OK(cudaMalloc(&s->exch, 2 * (size_t) N * sizeof(float4)));
        DFU_CUDA_OK(cudaMalloc(&s->st, 6 * (size_t) N * sizeof(float4)));
        DFU_CUDA_OK(cudaMalloc(&s->xw, 2 * 3 * (size_t) N * sizeof(unsigned long long)));
        DFU_CUDA_OK(cudaMalloc(&s->t4, (size_t) N * sizeof(float4)));
        // version 4 re-uses xw as its tagged exchange vector: cleared once, sequence numbers never repeat afterwards
        DFU_CUDA_OK(cudaMemsetAsync(s->xw, 0, 2 * 3 * (size_t) N * sizeof(unsigned long long), st));
        s->cap_rows = (size_t) N;
    }
    if (!s->pat_cursor) DFU_CUDA_OK(cudaMalloc(&s->pat_cursor, sizeof(int)));
    if (!s->pw) {  // version 4: per-CTA inboxes of the tagged partial sums, [2][P4_MAX_CTAS][P4_MAX_CTAS] float4
        DFU_CUDA_OK(cudaMalloc(&s->pw, P4_PQ_BYTES));
        DFU_CUDA_OK(cudaMemsetAsync(s->pw, 0, P4_PQ_BYTES, st));
    }
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


// north-star extension: Gauss-Newton on the point-to-plane SE(3) energy; one cooperative launch (kp_persistent), or one
// kernel per phase (no cooperative launch / DFU_SOLVER_PATH=multi)
int solve_p2plane(dfu_solver* s, cudaStream_t st) {
    const int N = s->N, P = s->P;
    const dfu_solver_params& prm = s->prm;
    DFU_REQUIRE(s->live_n != nullptr, DFU_ERR_INVALID, "the point-to-plane energy needs the live normals (initializeProblemInstance)");
    DFU_REQUIRE(s->lists_sorted, DFU_ERR_NOT_INIT, "set the energy before initializeProblemInstance");
    // scratch per point (floats): wn 8 | jac 48 | ent 64 | tpos 8 | svT 8 | e | sv | tk (8 bytes) -> 140
    // scratch per node: 7 vectors of 8 | X 12 | G 48 | Minv 36 | Gd 24 | ew 8 | L 21 | rslot (8 bytes) 2 -> 207   (orders keep the vector-loaded arrays aligned)
    const size_t need_pt = (size_t) std::max(P, 1) * 140, need_node = (size_t) N * 207;
    if (need_pt > s->p2p_cap_pt) {
        cudaFree(s->p2p_pt);
        s->p2p_pt = nullptr; s->p2p_cap_pt = 0;
        DFU_CUDA_OK(cudaMalloc(&s->p2p_pt, need_pt * sizeof(float)));
        s->p2p_cap_pt = need_pt;
    }
    if (need_node > s->p2p_cap_node) {
        cudaFree(s->p2p_node);
        s->p2p_node = nullptr; s->p2p_cap_node = 0;
        DFU_CUDA_OK(cudaMalloc(&s->p2p_node, need_node * sizeof(float)));
        s->p2p_cap_node = need_node;
    }
    P2PProblem pb{};
    pb.N = N; pb.P = P;
    pb.nbr = s->nbr; pb.wts = s->wts; pb.canon = s->canon_v; pb.live = s->live_v; pb.nrm = s->live_n;
    pb.tptr = s->tptr; pb.tv = s->tv; pb.nnbr = s->nnbr; pb.rin_ptr = s->rin_ptr; pb.rin = s->rin; pb.pos_w = s->wf->pos_w;
    pb.wreg2 = prm.lambda / ((float) N * 8.f);
    pb.tukey_offset = prm.tukey_offset; pb.psi_data = prm.psi_data;
    float* pp = s->p2p_pt;
    pb.wn = pp; pp += (size_t) P * 8;
    pb.jac = pp; pp += (size_t) P * 48;
    pb.ent = reinterpret_cast<float4*>(pp); pp += (size_t) P * 64;
    pb.tpos = reinterpret_cast<int*>(pp); pp += (size_t) P * 8;
    pb.svT = pp; pp += (size_t) P * 8;
    pb.e = pp; pp += P;
    pb.sv = pp; pp += P;
    pb.tk = reinterpret_cast<unsigned char*>(pp);  // 8P bytes
    pb.theta = s->theta;
    float* pn = s->p2p_node;
    const size_t vs = (size_t) N * P2P_VS;
    pb.b = pn; pb.x = pn + vs; pb.r = pn + 2 * vs; pb.z = pn + 3 * vs; pb.p = pn + 4 * vs; pb.q = pn + 5 * vs; pb.p2 = pn + 6 * vs;
    pn += 7 * vs;
    pb.X = pn; pn += (size_t) N * 12;
    s->p2p_X = pb.X;
    pb.G = pn; pn += (size_t) N * 48;
    pb.Minv = pn; pn += (size_t) N * 36;
    pb.Gd = pn; pn += (size_t) N * 24;
    pb.ew = pn; pn += (size_t) N * 8;
    pb.reg_mode = s->reg_mode; pb.psi_reg = prm.psi_reg;
    pb.L = pn; pn += (size_t) N * 21;
    pb.rslot = reinterpret_cast<unsigned char*>(pn);  // 8N bytes
    pb.part = s->part;
    const int nblk_n = div_up(N, TPB), nblk_p = std::max(1, std::min(div_up(P, TPB), MAX_PARTIALS)),
              nblk_e = std::min(div_up((long) N * 8, TPB), MAX_PARTIALS), nblk_w = std::min(div_up((long) N * 32, TPB), MAX_PARTIALS);
    DFU_REQUIRE(div_up(P, TPB) <= MAX_PARTIALS && div_up((long) N * 8, TPB) <= MAX_PARTIALS, DFU_ERR_UNSUPPORTED,
                "point-to-plane mode: at most 262144 points / 32768 nodes");
    const double tol2 = (double) prm.pcg_tol * (double) prm.pcg_tol;
    const char* force = getenv("DFU_SOLVER_PATH");
    if (s->coop_blocks_p2p > 0 && !(force && force[0] == 'm')) {
        P2PCtl ctl{prm.num_iter, prm.nonlinear_iter, prm.linear_iter, prm.early_out, tol2, 16, nullptr};
        if (const char* e = getenv("DFU_P2P_REFRESH")) ctl.refresh = std::max(1, atoi(e));  // experiments
        static long long* prof_dev = nullptr;
        const bool profile = getenv("DFU_SOLVER_PROFILE") != nullptr;
        if (profile) {
            if (!prof_dev) DFU_CUDA_OK(cudaMalloc(&prof_dev, P2P_PROF_N * sizeof(long long)));
            DFU_CUDA_OK(cudaMemsetAsync(prof_dev, 0, P2P_PROF_N * sizeof(long long), st));
            ctl.prof = prof_dev;
        }
        Scalars* sc = s->sc;
        unsigned* bar = s->bar;
        float4 *real = s->wf->real, *dual = s->wf->dual;
        int blocks = s->coop_blocks_p2p;
        if (const char* e = getenv("DFU_P2P_CTAS")) blocks = std::max(1, std::min(blocks, atoi(e)));  // experiments: fewer CTAs
        DFU_CUDA_OK(cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned), st));
        s->bar_clean = false;
        void* args[] = {&pb, &ctl, &sc, &bar, &real, &dual};
        DFU_CUDA_OK(cudaLaunchCooperativeKernel((void*) kp_persistent, dim3(blocks), dim3(P2P_TPB), args, 0, st));
        ++g_dfu_launches;
        s->last_kernel = 4;
        s->gn_steps_host = -1;  // read from the device scalars
        if (getenv("DFU_DEBUG")) fprintf(stderr, "[dfu] point-to-plane persistent kernel, %d CTAs\n", blocks);
        if (profile) {  // debugging aid: synchronises
            long long h[P2P_PROF_N];
            DFU_CUDA_OK(cudaMemcpyAsync(h, prof_dev, sizeof(h), cudaMemcpyDeviceToHost, st));
            DFU_CUDA_OK(cudaStreamSynchronize(st));
            static const char* names[P2P_PROF_N] = {"init", "linearise", "bar", "assemble", "bar", "scalars", "point_apply", "bar", "node_apply",
                                                    "bar", "sum+update", "bar", "sum", "expmap+bar", "final", "-"};
            fprintf(stderr, "[dfu] p2p cycles of CTA 0 (%d CTAs):", blocks);
            for (int i = 0; i < P2P_PROF_N - 1; ++i) fprintf(stderr, " %s=%lld", names[i], h[i]);
            fprintf(stderr, "\n");
        }
        return dfu_wf_refresh_flags(s->wf, st);
    }
    kp_init<<<std::max(nblk_n, div_up(P, TPB)), TPB, 0, st>>>(pb, s->sc);  // also resets the device scalars
    DFU_LAUNCH_OK();
    kp_slots<<<nblk_w, TPB, 0, st>>>(pb);
    DFU_LAUNCH_OK();
    bool stop_all = false;
    s->gn_steps_host = -1;
    for (int outer = 0; outer < prm.num_iter && !stop_all; ++outer) {
        for (int gn = 0; gn < prm.nonlinear_iter; ++gn) {
            kp_linearise<<<nblk_p, TPB, 0, st>>>(pb, gn == 0 ? 1 : 0);
            DFU_LAUNCH_OK();
            kp_edges<<<nblk_e, TPB, 0, st>>>(pb, gn == 0 ? 1 : 0);
            DFU_LAUNCH_OK();
            kp_assemble<<<nblk_w, TPB, 0, st>>>(pb);
            DFU_LAUNCH_OK();
            kp_init_scalars<<<1, TPB, 0, st>>>(pb, s->sc, nblk_p, nblk_e, nblk_w, tol2);
            DFU_LAUNCH_OK();
            if (prm.early_out) {
                int rc = read_scalars(s, st);
                if (rc != DFU_OK) return rc;
                if (s->sc_host->done_it == 0) {
                    if (gn == 0 && outer > 0) stop_all = true;
                    break;
                }
            }
            for (int it = 0; it < prm.linear_iter; ++it) {
                kp_point_apply<<<nblk_p, TPB, 0, st>>>(pb, s->sc, it);
                DFU_LAUNCH_OK();
                kp_node_apply<<<nblk_w, TPB, 0, st>>>(pb, s->sc, it);
                DFU_LAUNCH_OK();
                kp_update<<<nblk_n, TPB, 0, st>>>(pb, s->sc, it, nblk_w);
                DFU_LAUNCH_OK();
                kp_direction<<<nblk_n, TPB, 0, st>>>(pb, s->sc, it, nblk_n, nblk_w, tol2);
                DFU_LAUNCH_OK();
                if (prm.early_out && (it & 15) == 15) {
                    int rc = read_scalars(s, st);
                    if (rc != DFU_OK) return rc;
                    if (s->sc_host->done_it <= it + 1) break;
                }
            }
            kp_expmap<<<nblk_n, TPB, 0, st>>>(pb, s->sc);
            DFU_LAUNCH_OK();
        }
    }
    // energy at the solution (Tukey weights of the last outer iteration; at the identity if no step ran)
    kp_linearise<<<nblk_p, TPB, 0, st>>>(pb, prm.num_iter * prm.nonlinear_iter == 0 ? 1 : 0);
    DFU_LAUNCH_OK();
    kp_edges<<<nblk_e, TPB, 0, st>>>(pb, prm.num_iter * prm.nonlinear_iter == 0 ? 1 : 0);
    DFU_LAUNCH_OK();
    kp_final_energy<<<1, TPB, 0, st>>>(pb, s->sc, nblk_p, nblk_e);
    DFU_LAUNCH_OK();
    // compose the increments onto the nodes once, like the reference does with its translations (opt_solver.cpp:270-285)
    kp_compose<<<nblk_n, TPB, 0, st>>>(pb, s->wf->real, s->wf->dual);
    DFU_LAUNCH_OK();
    s->last_kernel = 0;
    return dfu_wf_refresh_flags(s->wf, st);
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
    if (e1 == cudaSuccess) e1 = cudaMemset(s->sc, 0, sizeof(Scalars));  // (spin_fail is only written by the kernels that can set it)
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
    {
        const int xs_max = (P4_MAX_N + 15) / 16 * 16 * (int) sizeof(float4);
        if (coop && cudaFuncSetAttribute(k_solve_persistent4, cudaFuncAttributeMaxDynamicSharedMemorySize, xs_max) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_persistent4, PTPB, xs_max) == cudaSuccess && per_sm >= 1)
            s->coop_blocks4 = min(sms, MAX_PARTIALS);
    }
    if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kp_persistent, P2P_TPB, 0) == cudaSuccess && per_sm >= 1)
        s->coop_blocks_p2p = min(sms * min(per_sm, P2P_CTAS_PER_SM), MAX_PARTIALS);
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
    cudaFree(s->p2p_pt);
    cudaFree(s->p2p_node);
    cudaFree(s->pat_cursor);
    cudaFree(s->sc);
    cudaFreeHost(s->sc_host);
    cudaFree(s->part);
    cudaFree(s->bar);
    cudaSetDevice(prev);
    delete s;
    return DFU_OK;
}

int dfu_solver_set_energy(dfu_solver* s, int energy_mode) {
    DFU_REQUIRE(s, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(energy_mode == DFU_ENERGY_REF_TRANSLATION || energy_mode == DFU_ENERGY_P2PLANE_SE3, DFU_ERR_INVALID, "bad energy mode");
    s->energy_mode = energy_mode;
    s->problem_ready = false;  // the transposed lists depend on the mode
    return DFU_OK;
}

int dfu_solver_set_regulariser(dfu_solver* s, int reg_mode) {
    DFU_REQUIRE(s, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(reg_mode == DFU_REG_QUADRATIC || reg_mode == DFU_REG_HUBER_ALPHA, DFU_ERR_INVALID, "bad regulariser");
    s->reg_mode = reg_mode;
    return DFU_OK;
}

int dfu_solver_get_increments(const dfu_solver* s, float* X12, dfu_stream stream) {
    DFU_REQUIRE(s && X12, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(s->energy_mode == DFU_ENERGY_P2PLANE_SE3 && s->p2p_X, DFU_ERR_NOT_INIT, "no point-to-plane solve has run");
    DFU_CUDA_OK(cudaMemcpyAsync(X12, s->p2p_X, 12 * (size_t) s->N * sizeof(float), cudaMemcpyDeviceToDevice, as_stream(stream)));
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
    (void) canon_n; (void) affine_host;  // uploaded but never read by the reference's energy (live_n: point-to-plane mode only)
    DFU_REQUIRE(s, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(P >= 0 && (P == 0 || (canon_v && live_v)), DFU_ERR_INVALID, "bad point arrays");
    dfu_warpfield* wf = s->wf;
    DFU_REQUIRE(wf->initialised, DFU_ERR_NOT_INIT, "warp field not initialised");
    DFU_REQUIRE(wf->N >= DFU_KNN, DFU_ERR_PRECONDITION, "the solver needs at least 8 nodes (reference UB, opt_solver.cpp:63-66)");
    DFU_REQUIRE((long) P * 8 < 0x7fffffffL, DFU_ERR_INVALID, "too many points");
    DFU_GUARD(wf->device);  // (restores the caller's device on every return path; drops stale errors of other libraries)
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
    s->canon_v = canon_v; s->live_v = live_v; s->live_n = live_n;
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
    s->lists_sorted = !pattern_eligible(s);
    if (P > 0) {
        // (unsorted lists: the counting atomics of the kNN kernel hand every edge its position in its node's list -- s->tent is
        //  free to hold them -- and the fill below is a plain scatter)
        rc = dfu_wf_build_data_graph(wf, canon_v, live_v, P, s->nbr, s->wts, s->dvec, s->tmp, s->lists_sorted ? nullptr : s->tent, st);
        if (rc != DFU_OK) return rc;
    }
    k_scan<<<1, 1024, 0, st>>>(s->tmp, N, s->tptr);
    DFU_LAUNCH_OK();
    if (P > 0) {
        if (s->lists_sorted) {
            k_fill<<<div_up(ne, TPB), TPB, 0, st>>>(s->nbr, ne, s->tptr, s->tmp + N, s->tent, 0);
            DFU_LAUNCH_OK();
            k_sort_emit<<<min(div_up((long) N * 32, TPB), 65535), TPB, 0, st>>>(s->tptr, N, s->tent, s->wts, s->tv, s->tw);
            DFU_LAUNCH_OK();
        } else {
            k_fill_emit_ranked<<<div_up(ne, TPB), TPB, 0, st>>>(s->nbr, s->tent, ne, s->tptr, s->wts, s->tv, s->tw);
            DFU_LAUNCH_OK();
        }
    }
    DFU_CUDA_OK(cudaMemsetAsync(s->vec, 0, 18 * (size_t) N * sizeof(float), st));  // unknowns := 0 (opt_solver.cpp:192-193)
    rc = build_pattern(s, st);
    if (rc != DFU_OK) return rc;
    s->problem_ready = true;
    return DFU_OK;
}

int dfu_solver_solve_all(dfu_solver* s, dfu_stream stream) {
    DFU_REQUIRE(s, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(s->problem_ready, DFU_ERR_NOT_INIT, "initializeProblemInstance has not been called");
    DFU_GUARD(s->wf->device);
    cudaStream_t st = as_stream(stream);
    // DFU_SOLVER_PATH=multi forces the one-kernel-per-phase path (used by the tests to cover both)
    const char* force = getenv("DFU_SOLVER_PATH");
    if (s->energy_mode == DFU_ENERGY_P2PLANE_SE3) {
        return solve_p2plane(s, st);
    }
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
    s->writeback_fused = false;
    int rc = multi ? solve_multi_kernel(s, st) : solve_persistent(s, st);
    if (rc != DFU_OK) return rc;
    // write back ONCE: dg_se3 := DQ(0,0,0,t) * dg_se3 (opt_solver.cpp:270-285, node.cpp:19-23) -- version 3r has done it
    if (s->writeback_fused) return DFU_OK;
    return dfu_warpfield_update_translations(s->wf, s->vec, stream);
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

namespace {
__global__ void k_export_stats(const Scalars* __restrict__ sc, int gn_steps_host, double* __restrict__ out) {
    out[0] = sc->E0;
    out[1] = sc->E;
    out[2] = (double) sc->pcg_iters;
    out[3] = (double) (gn_steps_host >= 0 ? gn_steps_host : sc->gn_steps);
}
}  // namespace

int dfu_solver_get_stats(const dfu_solver* s, double* stats_dev, dfu_stream stream) {
    DFU_REQUIRE(s && stats_dev, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(s->problem_ready, DFU_ERR_NOT_INIT, "no problem instance");
    (void) cudaGetLastError();
    k_export_stats<<<1, 1, 0, as_stream(stream)>>>(s->sc, s->gn_steps_host, stats_dev);
    DFU_LAUNCH_OK();
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
    DFU_REQUIRE(!(s->last_kernel == 5 && s->sc_host->spin_fail), DFU_ERR_CUDA,
                "persistent solver: an exchanged word never arrived (grid not co-resident?); result invalid");
    return DFU_OK;
}

}  // extern "C"
