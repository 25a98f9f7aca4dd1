/*
 * dynfu_oracle.h -- C interface of the CPU ORACLE for the dynfu hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The product
 * (libdynfu_b200.so) never links, loads or calls anything in oracle/.
 *
 * Every function restates, on the CPU and with no third-party dependency, the arithmetic of
 * a function of the reference swarth100/dynfu; the reference file:line each one follows is
 * given next to its definition in dynfu_oracle.cpp.
 *
 * Parity status:
 *   - dual-quaternion algebra : PINNED by the 23 golden cases of test/quaternion_test.cpp
 *   - kNN                     : PINNED against the reference's vendored nanoflann
 *                               (oracle/_ref/libdynfu_oracle_nf.so, compiled from the reference header)
 *   - blend / warp            : PINNED through the DQ goldens + the OptTest post-conditions
 *   - TSDF integrate          : parity UNPINNED -- the reference has no CPU integrator and no test
 *                               of it; this is a restatement of its CUDA kernel in a defined
 *                               IEEE arithmetic (see DESIGN.md "canonical arithmetic")
 *   - solver                  : parity UNPINNED at the Opt boundary (Opt/Terra are not in the tree,
 *                               Ceres is never linked); pinned only by the 8 OptTest post-conditions
 *   - 1-NN correspondences    : PINNED against the reference's vendored nanoflann (as kNN)
 *   - points/normals, raycast : parity UNPINNED -- no reference test; restated in IEEE arithmetic (division and
 *                               sqrt instead of the GPU's approximate intrinsics), checked against analytic scenes
 *   - Warpfield::update       : parity UNPINNED -- PCL 1.8.1's VoxelGrid is not vendored; its published algorithm
 *                               is restated, float additions inside a cell in ascending point index
 *   - point-to-plane SE(3)    : parity UNPINNED -- no reference implementation (north-star extension); pinned
 *                               against scipy.optimize.least_squares in tests/test_oracle_p2plane.py; its robust
 *                               regulariser (reg_mode 1: alpha_ij * Huber, the term opt_solver.cpp:233-268 and
 *                               energy.t:76 prepare and the reference leaves out) likewise
 */
#ifndef DYNFU_ORACLE_H
#define DYNFU_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* quaternions are (w,x,y,z); a dual quaternion is 8 floats: real wxyz then dual wxyz */

enum { ORC_BLEND_REF_COMPOSE = 0, ORC_BLEND_DQB_SUM = 1 };
enum { ORC_NORMAL_REF = 0, ORC_NORMAL_ROTATE_ONLY = 1 };

/* ---- dual quaternion algebra (include/dynfu/utils/dual_quaternion.hpp) ---- */
void orc_dq_from_rot_trans(const float rot[4], const float t[3], float out[8]);
void orc_dq_from_euler(float yaw, float pitch, float roll, float x, float y, float z, float out[8]);
void orc_dq_from_rodrigues(const float rod[3], const float t[3], float out[8]);
void orc_dq_add(const float a[8], const float b[8], float out[8]);
void orc_dq_sub(const float a[8], const float b[8], float out[8]);
void orc_dq_scale(const float a[8], float s, float out[8]);
void orc_dq_mul(const float a[8], const float b[8], float out[8]);
void orc_dq_conj(const float a[8], float out[8]);
int  orc_dq_normalize(const float a[8], float out[8]); /* returns 0 ok, 1 if |real| <= FLT_EPS */
void orc_dq_get_translation(const float a[8], float t[3]);
float orc_dq_get_roll(const float a[8]);
float orc_dq_get_pitch(const float a[8]);
float orc_dq_get_yaw(const float a[8]);
void orc_dq_get_rodrigues(const float a[8], float rod[3]);
void orc_dq_transform_vertex(const float a[8], const float v[3], float out[3]);
void orc_dq_transform_normal(const float a[8], const float n[3], int normal_mode, float out[3]);
/* writes "real: (w,x,y,z)\ndual: (w,x,y,z)\n" (operator<<); returns length */
int  orc_dq_to_string(const float a[8], char* buf, int buflen);

/* ---- node weight (src/dynfu/utils/node.cpp:29-36) ---- */
float orc_node_weight(const float node_pos[3], float dg_w, const float p[3]);

/* ---- kNN (src/dynfu/warp_field.cpp:111-122 + nanoflann metric) ----
 * brute force, key (dist2, idx) lexicographic, ascending.  k <= 16.  Returns the number of
 * queries whose (k+1) smallest distances contain a bit-equal pair (ties): must be 0 for the
 * result to be comparable with nanoflann's visitation-order tie break. */
long orc_knn(const float* nodes_xyz, int N, const float* q_xyz, long Q, int k,
             int32_t* idx_out, float* dist2_out_or_null);

/* ---- blend / warp (src/dynfu/warp_field.cpp:127-171) ---- */
/* nodes: pos_xyz[N*3], dq[N*8], dg_w[N].  k is fixed at 8 (KNN, warp_field.hpp:27). */
void orc_blend(const float* pos, const float* dq, const float* dg_w, int N,
               const float* p_xyz, long Q, int blend_mode, float* dq_out);
void orc_warp(const float* pos, const float* dq, const float* dg_w, int N,
              const float* v, const float* n_or_null, long P, int blend_mode, int normal_mode,
              float* v_out, float* n_out_or_null);

/* ---- depth -> ray length (src/kfusion/cuda/imgproc.cu:233-245) ---- */
void orc_compute_dists(const uint16_t* depth, size_t depth_pitch_bytes, uint16_t* dists,
                       size_t dists_pitch_bytes, int rows, int cols, const float intr[4]);

/* ---- half helpers (include/kfusion/cuda/device.hpp:59-67) ---- */
/* ---- depth -> points + normals (src/kfusion/cuda/imgproc.cu:187-215) ---- */
void orc_points_normals(const uint16_t* depth, size_t depth_pitch_bytes, int rows, int cols, const float intr[4],
                        float* points4, float* normals4);
long orc_compact_points(const float* points4, const float* normals4_or_null, int rows, int cols, const float* xform,
                        float* out_v, float* out_n_or_null);
/* ---- live -> canonical correspondences (src/dynfu/dyn_fusion.cpp:212-242) ---- */
long orc_find_corresponding(const float* canon_v, const float* canon_n_or_null, int P_canon, const float* live_v,
                            long P_live, float* out_v, float* out_n_or_null, int32_t* idx_out_or_null);
/* ---- Warpfield::update (src/dynfu/warp_field.cpp:34-95) ---- */
long orc_unsupported(const float* pos, const float* dg_w, int N, const float* verts, long P, uint8_t* flags_out);
long orc_voxel_grid(const float* pts, long U, const float leaf[3], float* out, int order_mode);
long orc_warpfield_update(const float* pos, const float* dq, const float* dg_w, int N, float epsilon,
                          const float* verts, long P, int blend_mode, float* pos_out, float* dq_out, float* w_out);
/* ---- TsdfVolume::raycast (src/kfusion/cuda/tsdf_volume.cu:126-386) ---- */
void orc_raycast(const uint32_t* vol, const int dims[3], const float voxel[3], float trunc, const float cam2vol[12],
                 const float rinv[9], const float intr[4], int rows, int cols, float step_factor, float grad_factor,
                 float* points4, float* normals4, uint16_t* depth_or_null);
uint16_t orc_float2half(float f);
float    orc_half2float(uint16_t h);

/* ---- TSDF (src/kfusion/cuda/tsdf_volume.cu:11-22, 43-94) ---- */
void orc_tsdf_clear(uint32_t* vol, const int dims[3], int z0, int z1);
/* nodes may be NULL (N=0) -> rigid.  Processes planes [z0,z1).  If f32_out (dims voxels of
 * float2 {tsdf, weight}) is non-NULL the pre-rounding float result of touched voxels is also
 * written there (untouched voxels are left as they were).  Returns #touched voxels. */
long orc_tsdf_integrate(uint32_t* vol, const int dims[3], const float voxel[3], float trunc,
                        int max_weight, const float vol2cam[12], const float intr[4],
                        const uint16_t* dists, size_t pitch_bytes, int rows, int cols,
                        const float* pos, const float* dq, const float* dg_w, int N,
                        int blend_mode, int z0, int z1, float* f32_out_or_null);
float orc_trunc_dist(float requested, const float voxel[3]); /* tsdf_volume.cpp:57-61 */

/* ---- robust weights (src/dynfu/utils/opt_solver.cpp:204-212, 233-239) ---- */
float orc_tukey(float tukey_offset, float c, const float err[3]);
float orc_huber(float k, float e);
void orc_huber_weights(const float* pos, const float* dq, int N, float psi_reg, float* out); /* opt_solver.cpp:241-268 */

/* ---- solver (opt_solver.cpp + energy.t; contract in DESIGN.md) ---- */
typedef struct {
    int   num_iter;        /* outer iterations (Tukey re-weighting)          */
    int   nonlinear_iter;  /* GN steps per outer iteration                   */
    int   linear_iter;     /* PCG steps per GN step                          */
    float tukey_offset, psi_data, lambda, psi_reg;
    double pcg_tol;        /* PCG stops when r.z <= tol^2 * (r.z of the first GN step); 0 = never */
    int   early_out;       /* stop outer loop when relative energy change < 1e-12 */
    int   reg_mode;        /* P2PLANE only. 0: w_reg^2 |X_i g_j - X_j g_j|^2 on every edge; 1: DynamicFusion eq. 8,
                              w_reg^2 alpha_ij psi_reg(.) with alpha_ij = max(dg_w_i, dg_w_j) and psi_reg = Huber, as IRLS:
                              edge weight alpha_ij h_ij, h re-evaluated with the Tukey weights (opt_solver.cpp:233-268) */
} orc_solver_params;

/* Solves energy.t (translation-only point-to-point, un-normalised Gaussian weights) in double.
 * t_out[N*3]: optimal translations (NOT yet composed onto the node DQs);
 * dq_inout[N*8]: node DQs, updated ONCE as DQ(0,0,0,t) * dq (node.cpp:19-23).
 * stats_out[4] = {initial energy, final energy, total PCG iterations, total GN steps}. */
int orc_solve(const float* pos, float* dq_inout, const float* dg_w, int N,
              const float* canon, const float* live, long P,
              const orc_solver_params* prm, double* t_out, double* stats_out);
/* energy of energy.t at translations t (double) with Tukey weights computed at t_tukey */
double orc_energy(const float* pos, const float* dg_w, int N, const float* canon,
                  const float* live, long P, const orc_solver_params* prm,
                  const double* t, const double* t_tukey);

/* North-star extension P2PLANE_SE3 (no reference implementation; see the .cpp): point-to-plane data term, one rigid
 * increment X_k per node (X_out[N*12]: R row-major then t), composed onto dq_inout once at the end. */
int orc_solve_p2plane(const float* pos, float* dq_inout, const float* dg_w, int N, const float* canon, const float* live,
                      const float* live_n, long P, const orc_solver_params* prm, double* X_out, double* stats_out);
double orc_energy_p2plane(const float* pos, const float* dg_w, int N, const float* canon, const float* live,
                          const float* live_n, long P, const orc_solver_params* prm, const double* X, const double* X_tukey);

/* ---- marching cubes (src/kfusion/cuda/marching_cubes.cu:33-260) : vertices of the zero level set, fixed tile order ---- */
long orc_marching_cubes(const uint32_t* vol, const int dims[3], const float volume_size[3], const signed char* tri, float* verts4,
                        int32_t* cube_ids, long capacity);

void orc_set_num_threads(int n);
int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
