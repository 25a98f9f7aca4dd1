/*
 * dynfu_b200.h -- C-ABI of the B200-native DynamicFusion hot path (libdynfu_b200.so, sm_100a).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Each entry point names
 * the interface of the reference (swarth100/dynfu, paths relative to its root) that it replaces.
 * INTEGRATION.md shows the reference-side binding a maintainer would add.
 *
 * Conventions
 *   - every function returns a dfu_status (0 = ok); nothing throws or exit()s across the boundary
 *     (the reference prints and exit(0)s on CUDA errors, include/kfusion/safe_call.hpp:12-22);
 *     dfu_last_error() gives the message of the calling thread's last failure;
 *   - pointers are DEVICE pointers unless the parameter or function name ends in _host; small POD
 *     parameter blocks (dims, voxel size, intrinsics, vol2cam, params structs) are read on the host
 *     at call time, like the by-value POD views of include/kfusion/internal.hpp:36-55;
 *   - every call takes a cudaStream_t (as void*), is asynchronous with respect to the host and adds
 *     no hidden device synchronisation (the reference cudaDeviceSynchronize()s after integrate,
 *     src/kfusion/cuda/tsdf_volume.cu:120); *_host variants synchronise the stream before returning;
 *   - the caller owns every buffer; handles own only their internal scratch; the TSDF volume is
 *     borrowed (it is the reference's own ushort2 blob, src/kfusion/tsdf_volume.cpp:32-38);
 *   - quaternions are (w,x,y,z), Hamilton product; a dual quaternion is 8 floats, real then dual
 *     (include/dynfu/utils/dual_quaternion.hpp); points/normals are packed xyz floats;
 *   - k is fixed at 8 (KNN, include/dynfu/warp_field.hpp:27).
 *
 * There is NO CPU fallback anywhere behind this header: without a CUDA device every compute entry
 * point fails with DFU_ERR_CUDA.
 */
#ifndef DYNFU_B200_H
#define DYNFU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFU_KNN 8
#define DFU_VERSION 100

typedef enum {
    DFU_OK = 0,
    DFU_ERR_INVALID = 1,      /* bad argument (null pointer, bad size, unsupported dims)           */
    DFU_ERR_CUDA = 2,         /* CUDA runtime error; message in dfu_last_error()                    */
    DFU_ERR_PRECONDITION = 3, /* e.g. fewer than 8 nodes for the solver (reference UB,              */
                              /* src/dynfu/utils/opt_solver.cpp:63-66)                              */
    DFU_ERR_NOT_INIT = 4,     /* handle used before init (nanoflann throws, nanoflann.hpp:1209)     */
    DFU_ERR_UNSUPPORTED = 5   /* input outside what the device tables hold (message says which)     */
} dfu_status;

/* how the 8 weighted node transforms are combined */
typedef enum {
    DFU_BLEND_REF_COMPOSE = 0, /* what Warpfield::calcDQB does (src/dynfu/warp_field.cpp:127-148):   */
                               /* Q = prod_k (real_k, w_k*dual_k), nearest first, real normalised    */
    DFU_BLEND_DQB_SUM = 1      /* true dual-quaternion blending: Q = sum_k w_k q_k / |real|          */
} dfu_blend_mode;

typedef enum {
    DFU_NORMAL_REF = 0,        /* DualQuaternion::transformNormal (dual_quaternion.hpp:217-228):     */
                               /* the vertex formula, translation included                           */
    DFU_NORMAL_ROTATE_ONLY = 1
} dfu_normal_mode;

typedef struct dfu_warpfield dfu_warpfield; /* replaces class Warpfield (include/dynfu/warp_field.hpp:32-78) */
typedef struct dfu_solver dfu_solver;       /* replaces class CombinedSolver (include/dynfu/utils/opt_solver.hpp:19-110) */
typedef void* dfu_stream;                   /* cudaStream_t */

int dfu_version(void);
const char* dfu_last_error(void);
/* number of CUDA kernels this library has launched so far in this process (instrumentation) */
unsigned long long dfu_launch_count(void);
/* 0 if a usable sm_100 device is visible, DFU_ERR_CUDA otherwise */
int dfu_device_check(int device);
/* Measurement aid (no reference counterpart; SURVEY.md section 6 asks for the FP32 SIMT peak as the denominator of the
 * kNN roofline): times an FFMA-chain kernel on `device` (synchronous) and returns the achieved TFLOP/s and the SM clock
 * that rate implies (MHz, may be NULL). */
int dfu_microbench_fp32(int device, double* tflops_out, double* sm_mhz_out);

/* ------------------------------------------------------------------------------------------------
 * Warp field  (replaces Warpfield + Node + the nanoflann KD-tree + DualQuaternion on the hot path)
 * ---------------------------------------------------------------------------------------------- */

/* Warpfield::Warpfield() (src/dynfu/warp_field.cpp:6) */
int dfu_warpfield_create(dfu_warpfield** out, int device);
int dfu_warpfield_destroy(dfu_warpfield* wf);

/* Warpfield::init(epsilon, nodes) (src/dynfu/warp_field.cpp:10-28): stores the nodes and builds the
 * search structure.  pos_xyz[N*3] = Node::dg_v, dq[N*8] = Node::dg_se3, dg_w[N] = Node::dg_w
 * (include/dynfu/utils/node.hpp:55-58).  N >= 1. */
int dfu_warpfield_init(dfu_warpfield* wf, float epsilon, const float* pos_xyz, const float* dq,
                       const float* dg_w, int N, dfu_stream stream);
int dfu_warpfield_init_host(dfu_warpfield* wf, float epsilon, const float* pos_xyz_host,
                            const float* dq_host, const float* dg_w_host, int N, dfu_stream stream);

/* Warpfield::getNodes() (src/dynfu/warp_field.cpp:32): any output pointer may be NULL */
int dfu_warpfield_num_nodes(const dfu_warpfield* wf, int* N_host);
int dfu_warpfield_get_nodes(const dfu_warpfield* wf, float* pos_xyz, float* dq, float* dg_w, dfu_stream stream);
int dfu_warpfield_get_nodes_host(const dfu_warpfield* wf, float* pos_xyz_host, float* dq_host,
                                 float* dg_w_host, dfu_stream stream);

/* Node::setTransformation for all nodes (src/dynfu/utils/node.cpp:25) */
int dfu_warpfield_set_transforms(dfu_warpfield* wf, const float* dq, dfu_stream stream);
int dfu_warpfield_set_transforms_host(dfu_warpfield* wf, const float* dq_host, dfu_stream stream);
/* Node::updateTransformation with DQ(0,0,0,t) for all nodes: dg_se3 := DQ(0,0,0,t_i) * dg_se3
 * (src/dynfu/utils/node.cpp:19-23 as called from opt_solver.cpp:270-285).  t_xyz[N*3]. */
int dfu_warpfield_update_translations(dfu_warpfield* wf, const float* t_xyz, dfu_stream stream);

/* Warpfield::findNeighborsIndex(8, vertex) for Q query points (src/dynfu/warp_field.cpp:111-122).
 * idx[Q*8] ascending by (squared distance, index); dist2 may be NULL.  Squared distances are the
 * float ((dx*dx)+dy*dy)+dz*dz of nanoflann's L2_Simple_Adaptor (nanoflann.hpp:338-345), bit for bit.
 * With fewer than 8 nodes the missing entries are -1 / +inf (the reference returns a shorter vector). */
int dfu_warpfield_knn(const dfu_warpfield* wf, const float* q_xyz, int Q, int32_t* idx, float* dist2,
                      dfu_stream stream);

/* Warpfield::calcDQB(point) for Q points (src/dynfu/warp_field.cpp:127-148): dq_out[Q*8] */
int dfu_warpfield_blend(const dfu_warpfield* wf, const float* p_xyz, int Q, float* dq_out, int blend_mode,
                        dfu_stream stream);

/* Warpfield::warpToLive(frame) (src/dynfu/warp_field.cpp:150-171): v_out[i] = calcDQB(v[i]).transformVertex(v[i]),
 * n_out[i] = calcDQB(v[i]).transformNormal(n[i]).  n / n_out may both be NULL.  In-place is allowed. */
int dfu_warpfield_warp(const dfu_warpfield* wf, const float* v_xyz, const float* n_xyz, int P, float* v_out,
                       float* n_out, int blend_mode, int normal_mode, dfu_stream stream);

/* Warpfield::warpToLive for a point set that does not change between calls -- the canonical frame, which DynFusion
 * warps again every frame (src/dynfu/dyn_fusion.cpp:196).  The 8 nearest nodes of a point and their weights depend
 * only on the point and the node POSITIONS (never on the node transforms), so they are computed on the first call
 * and kept in `cache` until the points or the node positions change: `points_version` is the caller's change
 * counter for v_xyz (same pointer + same version + same P = same points); node positions are tracked by the
 * library.  Results are bit-identical to dfu_warpfield_warp.  Not in place (v_out != v_xyz). */
typedef struct dfu_pointcache dfu_pointcache;
int dfu_pointcache_create(dfu_pointcache** out, int device);
int dfu_pointcache_destroy(dfu_pointcache* cache);
int dfu_warpfield_warp_cached(const dfu_warpfield* wf, dfu_pointcache* cache, unsigned long long points_version,
                              const float* v_xyz, const float* n_xyz, int P, float* v_out, float* n_out,
                              int blend_mode, int normal_mode, dfu_stream stream);

/* ------------------------------------------------------------------------------------------------
 * TSDF volume  (layout of kfusion::cuda::TsdfVolume: ushort2 {half tsdf bits, u16 weight} per voxel,
 * idx = x + y*dims.x + z*dims.x*dims.y, include/kfusion/cuda/device.hpp:20-35,59-67)
 * ---------------------------------------------------------------------------------------------- */

/* cuda::computeDists (src/kfusion/imgproc.cpp:38-41 -> src/kfusion/cuda/imgproc.cu:233-254):
 * depth u16 millimetres -> ray length in metres as half bits.  intr_host = {fx, fy, cx, cy}. */
int dfu_compute_dists(const uint16_t* depth, size_t depth_pitch_bytes, uint16_t* dists, size_t dists_pitch_bytes,
                      int rows, int cols, const float intr_host[4], dfu_stream stream);

/* TsdfVolume::setTruncDist clamp (src/kfusion/tsdf_volume.cpp:57-61) */
float dfu_tsdf_trunc_dist(float requested, const float voxel_size_host[3]);

/* TsdfVolume::clear (src/kfusion/tsdf_volume.cpp:74-80 -> tsdf_volume.cu:11-34), planes [z0,z1) */
int dfu_tsdf_clear(void* volume, const int dims_host[3], int z0, int z1, dfu_stream stream);

/* TsdfVolume::integrate (src/kfusion/tsdf_volume.cpp:82-93 -> tsdf_volume.cu:43-121), planes [z0,z1).
 * vol2cam_host = 9 floats row-major R then 3 floats t (device::Aff3f, include/kfusion/internal.hpp:28-34)
 * = camera_pose.inv() * volume_pose.  wf == NULL integrates rigidly, exactly like the reference; with a
 * warp field every voxel position (volume-local metres, the frame of the nodes) is first moved by
 * calcDQB(p).transformVertex(p).  dims.x % 32 == 0 (src/kfusion/kinfu.cpp:47) and dims.y % 8 == 0. */
int dfu_tsdf_integrate(void* volume, const int dims_host[3], const float voxel_size_host[3], float trunc_dist,
                       int max_weight, const float vol2cam_host[12], const float intr_host[4],
                       const uint16_t* dists, size_t dists_pitch_bytes, int rows, int cols, dfu_warpfield* wf,
                       int blend_mode, int z0, int z1, dfu_stream stream);
/* The frame operator as one call: DynFusion::operator() (src/dynfu/dyn_fusion.cpp:48-145) -> warpCanonicalToLiveOpt
 * (:182-210) on device-resident inputs.  computeDists (:55) -> warpToLive(canonical) (:196) -> initializeProblemInstance
 * (:206) -> solveAll (:207; the node transforms are updated) -> integration of the live depth into the volume through the
 * solved field.  wf == NULL or solver == NULL or P == 0: frame 0, rigid integration only (:70).
 *   depth_mm      uint16 millimetres, row-pitched                 (cuda::Depth, include/kfusion/types.hpp)
 *   dists         scratch of the same shape for the ray lengths   (KinFu::dists_, include/kfusion/kinfu.hpp:104)
 *   canon_v       the canonical vertices [P][3]; canon_version = the caller's change counter for them (dfu_warpfield_warp_cached)
 *   canon_warped  scratch [P][3]: the canonical vertices warped to the previous live frame (getCanonicalWarpedToLive)
 *   live_v        the live vertices [P][3], paired with canon_v
 * Asynchronous on `stream`. */
typedef struct dfu_frame_params {
    void* volume;          /* ushort2 voxels, borrowed */
    int dims[3];
    float voxel_size[3];
    float trunc_dist;
    int max_weight;
    float vol2cam[12];     /* camera_pose.inv() * volume_pose: 9 floats row-major R, then t */
    float intr[4];         /* fx, fy, cx, cy */
    int rows, cols;
    int blend_mode;        /* DFU_BLEND_* */
    int z0, z1;            /* z-slab of this rank */
} dfu_frame_params;
int dfu_frame(dfu_warpfield* wf, dfu_solver* solver, dfu_pointcache* canon_cache, const dfu_frame_params* params,
              const uint16_t* depth_mm, size_t depth_pitch_bytes, uint16_t* dists, size_t dists_pitch_bytes, const float* canon_v,
              unsigned long long canon_version, float* canon_warped, const float* live_v, int P, dfu_stream stream);

/* kfusion::cuda::MarchingCubes::run (include/kfusion/cuda/marching_cubes.hpp:35-36, src/kfusion/marching_cubes.cpp:20-63 ->
 * device::getOccupiedVoxels / computeOffsetsAndTotalVertices / generateTriangles, src/kfusion/cuda/marching_cubes.cu:144-296):
 * triangle vertices of the zero level set of the volume as (x, y, z, 1) in volume-local metres, three consecutive vertices
 * per triangle -- what DynFusion::operator() downloads as the canonical / live surface points (dyn_fusion.cpp:74-88,120-134).
 * Any dims with dims[0] % 4 == 0 (the reference hard-codes 128^3); volume_size_host = the volume's edge lengths in metres
 * (cells are volume_size / dims, vertices carry the reference's half-cell shift); fixed output order (tiles of 32 x 8 x 8
 * cubes ascending, cubes ascending inside a tile) instead of the reference's atomic order.  Writes at most `capacity`
 * vertices (and, when cube_ids is non-NULL, the linear index x + dx (y + dy z) of each vertex's cube); *n_vertices_dev
 * (device) receives the number of vertices that exist, which may exceed capacity.  Asynchronous. */
int dfu_marching_cubes(const void* volume, const int dims[3], const float volume_size_host[3], void* vertices_xyz1,
                       int32_t* cube_ids, long capacity, int* n_vertices_dev, dfu_stream stream);

/* Instrumentation (no reference counterpart): counters of the LAST dfu_tsdf_integrate call on the current device, read back
 * synchronously: [0] voxels updated (tsdf_volume.cu:83-90 executed), [1] 16-byte quads read + written, [2] voxels among [0] updated through the saturated-free-space path (tsdf == 1 proven per brick),
 * [3] 8^3 bricks that ran the per-voxel warp.  The algorithmic TSDF traffic of the call is 8 B x stats[0]. */
int dfu_tsdf_integrate_stats(unsigned long long stats_host[4], dfu_stream stream);


/* TsdfVolume::raycast (src/kfusion/tsdf_volume.cpp:95-129 -> device::raycast, include/kfusion/internal.hpp, src/kfusion/
 * cuda/tsdf_volume.cu:126-386): first + -> - zero crossing of the TSDF along every pixel's ray, refined by trilinear
 * interpolation, normal from central differences of the interpolated TSDF.  cam2vol_host = volume_pose.inv() *
 * camera_pose as {R row-major, t}; rinv_host = the inverse of its rotation (the reference passes both, :98-102).
 * Outputs in the CAMERA frame like the reference: points4 and/or depth (u16 mm), normals4; NaN / 0 where a ray finds
 * no surface.  Needs the whole volume (not a z-slab).  Arithmetic: IEEE, one rounding per operation (see raycast.cu). */
int dfu_tsdf_raycast(const void* volume, const int dims_host[3], const float voxel_size_host[3], float trunc_dist,
                     const float cam2vol_host[12], const float rinv_host[9], const float intr_host[4], int rows, int cols,
                     float raycast_step_factor, float gradient_delta_factor, float* points4, size_t points_pitch_bytes,
                     uint16_t* depth, size_t depth_pitch_bytes, float* normals4, size_t normals_pitch_bytes,
                     dfu_stream stream);

/* ------------------------------------------------------------------------------------------------
 * Solver  (replaces CombinedSolver + Opt + energy.t)
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
    int num_iter;       /* CombinedSolverParameters::numIter      : outer iterations (Tukey re-weighting) */
    int nonlinear_iter; /* CombinedSolverParameters::nonLinearIter : Gauss-Newton steps per outer iteration */
    int linear_iter;    /* CombinedSolverParameters::linearIter    : PCG steps per Gauss-Newton step        */
    float tukey_offset; /* CombinedSolver ctor, src/dynfu/utils/opt_solver.cpp:3-13                        */
    float psi_data;
    float lambda;
    float psi_reg;
    float pcg_tol;      /* PCG stops when r.z <= tol^2 * (r.z of the first GN step); 0 = never             */
    int early_out;      /* CombinedSolverParameters::earlyOut                                               */
} dfu_solver_params;

/* all-reduce hook for data-parallel solves: sums buf[count] floats over ranks, in place, on stream */
/* Energy minimised by the solver.
 *   DFU_ENERGY_REF_TRANSLATION  the reference's energy.t: point-to-point, translations only (default; parity with the oracle)
 *   DFU_ENERGY_P2PLANE_SE3      north-star extension without a reference implementation: sum theta (n_live . (p - live))^2
 *                               with p = sum_k w^_k X_k canon (normalised weights), one rigid increment X_k per node,
 *                               as-rigid-as-possible regulariser w_reg^2 |X_n g_m - X_m g_m|^2; needs live_n in
 *                               dfu_solver_init_problem, whose point arrays must then stay valid until solve_all returns.
 *                               The increments are composed onto the node transforms once at the end. */
enum { DFU_ENERGY_REF_TRANSLATION = 0, DFU_ENERGY_P2PLANE_SE3 = 1 };

typedef int (*dfu_allreduce_fn)(float* buf, size_t count, void* ctx, dfu_stream stream);

/* CombinedSolver::CombinedSolver(warpfield, params, tukeyOffset, psi_data, lambda, psi_reg)
 * (src/dynfu/utils/opt_solver.cpp:3-13).  The solver SHARES the warp field's nodes, as the
 * reference's by-value Warpfield copy shares its shared_ptr<Node>s. */
int dfu_solver_create(dfu_solver** out, dfu_warpfield* wf, const dfu_solver_params* params_host);
int dfu_solver_destroy(dfu_solver* s);
/* ranks hold disjoint point partitions; fn sums the per-node normal-equation buffers over ranks */
int dfu_solver_set_allreduce(dfu_solver* s, dfu_allreduce_fn fn, void* ctx);

/* NCCL communicator owned by the library (libnccl.so.2 is opened at run time; nothing to link).  Rank 0 calls
 * dfu_comm_unique_id and ships the 128 bytes to the other ranks (e.g. through torch.distributed); every rank then
 * calls dfu_comm_create with its CUDA device current.  dfu_solver_set_comm makes the solver issue ncclAllReduce on
 * its own stream instead of calling a host callback; dfu_comm_allreduce is that callback (a dfu_allreduce_fn with
 * ctx = the dfu_comm). */
typedef struct dfu_comm dfu_comm;
int dfu_comm_unique_id(char id_host[128]);
int dfu_comm_create(dfu_comm** out, const char id_host[128], int rank, int world);
int dfu_comm_destroy(dfu_comm* c);
int dfu_comm_allreduce(float* buf, size_t count, void* ctx, dfu_stream stream);
int dfu_solver_set_comm(dfu_solver* s, dfu_comm* c);

/* CombinedSolver::initializeProblemInstance(canonicalFrame, liveFrame, affine)
 * (src/dynfu/utils/opt_solver.cpp:15-54): uploads nothing (pointers are device), builds the kNN data
 * graph (:56-72) and regularisation graph (:74-105), zeroes the unknowns (:192-193).
 * normals and affine may be NULL (the reference's energy never reads them, energy.t:29,32). */
int dfu_solver_init_problem(dfu_solver* s, const float* canon_v, const float* canon_n, const float* live_v,
                            const float* live_n, int P, const float affine_host[12], dfu_stream stream);

/* CombinedSolverBase::solveAll() [Opt, not in the reference tree]: runs the outer/GN/PCG loops on the
 * device and composes the result onto the warp field's nodes ONCE (opt_solver.cpp:270-285). */
int dfu_solver_solve_all(dfu_solver* s, dfu_stream stream);

/* results of the last solve_all: t_xyz[N*3] (device, may be NULL) and
 * stats_host[4] = {initial energy, final energy, PCG iterations, GN steps} (synchronises the stream) */
int dfu_solver_get_translations(const dfu_solver* s, float* t_xyz, dfu_stream stream);
/* call before dfu_solver_init_problem */
int dfu_solver_set_energy(dfu_solver* s, int energy_mode);
/* Regulariser of DFU_ENERGY_P2PLANE_SE3 (the reference energy ignores it), effective from the next solve_all:
 *   DFU_REG_QUADRATIC    w_reg^2 |X_i g_j - X_j g_j|^2 on every edge of the node graph -- the form energy.t:73-78 gives its
 *                        translation-only regulariser (default);
 *   DFU_REG_HUBER_ALPHA  the term the reference prepares and then leaves out (alpha: energy.t:76, Huber weights:
 *                        opt_solver.cpp:233-268, TODO at energy.t:2), i.e. DynamicFusion eq. 8:
 *                        w_reg^2 alpha_ij psi_reg(X_i g_j - X_j g_j), alpha_ij = max(dg_w_i, dg_w_j), psi_reg = Huber with
 *                        threshold dfu_solver_params.psi_reg -- solved as IRLS, the Huber weight of an edge re-evaluated
 *                        whenever the Tukey weights are (once per outer iteration). */
enum { DFU_REG_QUADRATIC = 0, DFU_REG_HUBER_ALPHA = 1 };
int dfu_solver_set_regulariser(dfu_solver* s, int reg_mode);
/* DFU_ENERGY_P2PLANE_SE3: the rigid increments of the last solve, N x 12 floats (R row-major, t) */
int dfu_solver_get_increments(const dfu_solver* s, float* X12, dfu_stream stream);
int dfu_solver_get_stats_host(const dfu_solver* s, double stats_host[4], dfu_stream stream);
/* the same four numbers written to DEVICE memory, stream-ordered and without a synchronisation (pipelined frame loops
 * copy them to pinned host memory together with the node transforms) */
int dfu_solver_get_stats(const dfu_solver* s, double* stats_dev, dfu_stream stream);

/* CombinedSolver::updateHuberWeights (opt_solver.cpp:241-268): huber[N], the value the reference's loop leaves
 * behind (its last neighbour); computed but never read by the reference's energy.  tukey[P] are the
 * calcTukeyBiweight values (opt_solver.cpp:204-231) the last solve ended with. */
int dfu_solver_huber_weights(const dfu_solver* s, float* huber, dfu_stream stream);
int dfu_solver_tukey_weights(const dfu_solver* s, float* tukey, dfu_stream stream);

/* ------------------------------------------------------------------------------------------------
 * Front end of the frame loop (the callers either side of the hot path)
 * ---------------------------------------------------------------------------------------------- */

/* cuda::computePointNormals (src/kfusion/imgproc.cpp:27-36 -> src/kfusion/cuda/imgproc.cu:187-226): depth u16 mm
 * -> camera-space points and normals, float4 images (x,y,z,0), all-NaN where a pixel or its right/lower neighbour
 * has no depth and in the last row/column.  Normalisation is cross / sqrt(dot) in IEEE arithmetic (the reference
 * uses the approximate rsqrt; difference <= 2 ulp). */
int dfu_compute_points_normals(const uint16_t* depth, size_t depth_pitch_bytes, int rows, int cols,
                               const float intr_host[4], float* points4, size_t points_pitch_bytes, float* normals4,
                               size_t normals_pitch_bytes, dfu_stream stream);

/* Packs the valid (non-NaN) pixels of a points image (and, if given, normals image) into xyz arrays in raster
 * order -- the device-side equivalent of downloading the cloud and pushing the valid entries into a
 * dynfu::Frame (src/dynfu/dyn_fusion.cpp:120-134, src/dynfu/utils/frame.cpp:3-18).  xform_host (NULL = none) is
 * a rigid transform {R row-major, t} applied to the points (rotation only to the normals), e.g. camera ->
 * volume-local.  *count_out (device int) receives the number of valid pixels; at most `capacity` are written. */
int dfu_compact_points(const float* points4, size_t points_pitch_bytes, const float* normals4,
                       size_t normals_pitch_bytes, int rows, int cols, const float xform_host[12], float* out_v,
                       float* out_n, int capacity, int* count_out, dfu_stream stream);

/* Exact nearest-neighbour index over an arbitrary point set: replaces the nanoflann KD-tree that
 * DynFusion::findCorrespondingFrame builds per frame (src/dynfu/dyn_fusion.cpp:212-242).  Ties resolve to the
 * lower index; distances use nanoflann's L2_Simple_Adaptor accumulation order. */
typedef struct dfu_pointindex dfu_pointindex;
int dfu_pointindex_create(dfu_pointindex** out, int device);
int dfu_pointindex_destroy(dfu_pointindex* pi);
int dfu_pointindex_build(dfu_pointindex* pi, const float* pts_xyz, int P, dfu_stream stream);
int dfu_pointindex_nearest(const dfu_pointindex* pi, const float* q_xyz, int Q, int32_t* idx, float* dist2,
                           dfu_stream stream);
/* findCorrespondingFrame in one call: out_v[i] / out_n[i] = canonical vertex / normal nearest to live_v[i].
 * canon_n, out_n, idx_out may be NULL. */
int dfu_find_corresponding(dfu_pointindex* pi, const float* canon_v, const float* canon_n, int P_canon,
                           const float* live_v, int P_live, float* out_v, float* out_n, int32_t* idx_out,
                           dfu_stream stream);

/* Warpfield::getUnsupportedVertices (src/dynfu/warp_field.cpp:34-62): flags[i] = 1 when
 * min_k |v_i - n_k| / dg_w_k >= 1 over the 8 nearest nodes (distance in double, stored as float, like the
 * reference's sqrt(pow()+pow()+pow())). */
int dfu_warpfield_unsupported(const dfu_warpfield* wf, const float* verts_xyz, int P, uint8_t* flags,
                              dfu_stream stream);

/* pcl::VoxelGrid<pcl::PointXYZ> with a cubic leaf (PCL 1.8.1, filters/impl/voxel_grid.hpp:212-437; the filter
 * Warpfield::update applies with a 5 cm leaf, src/dynfu/warp_field.cpp:68-72): one centroid per non-empty cell,
 * cells in ascending linear index.  Inside a cell the points are added in ascending index (PCL's std::sort leaves
 * that order implementation defined).  out_xyz needs room for U points; *M_host receives the number written.
 * Synchronises the stream.  DFU_ERR_UNSUPPORTED if the cell table would exceed 2^21 cells. */
int dfu_voxel_grid_filter(const float* pts_xyz, int U, float leaf, float* out_xyz, int* M_host, dfu_stream stream);

/* Warpfield::update (src/dynfu/warp_field.cpp:64-95): finds the unsupported vertices of the frame, decimates them
 * on a 5 cm grid, appends one node per centroid (dg_se3 = calcDQB(centroid) w.r.t. the nodes before the call,
 * dg_w = 2*epsilon) and re-indexes.  Existing node indices are preserved.  The integrator's per-voxel neighbour
 * cache is invalidated only for bricks a new node can reach.  Synchronises the stream (the node count changes). */
int dfu_warpfield_update(dfu_warpfield* wf, const float* verts_xyz, int P, int blend_mode, int* num_unsupported_host,
                         int* num_new_host, dfu_stream stream);

/* Diagnostics of the integrator's per-voxel neighbour cache: bricks the pool holds, bricks currently valid. */
int dfu_warpfield_cache_stats(const dfu_warpfield* wf, long long* pool_bricks_host, long long* built_bricks_host,
                              dfu_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* DYNFU_B200_H */
