// dfu_frame: the per-frame hot path as ONE C-ABI call ("dfu_frame(...) = the KinFu / DynFusion frame
// operator").  Host-side composition of the library's own entry points in the order of DynFusion::operator()
// (src/dynfu/dyn_fusion.cpp:48-145) -> warpCanonicalToLiveOpt (:182-210):
//   computeDists (:55) -> warpToLive(canonical) (:196) -> initializeProblemInstance (:206) -> solveAll (:207, writes the node
//   transforms) -> integration of the live depth through the solved field into the canonical volume (the reference's "future
//   addition", README.md:14-17; with warpfield == NULL the rigid TsdfVolume::integrate of :70).
// Everything is asynchronous on `stream`; no kernel of its own.
#include "dfu_internal.h"

extern "C" int dfu_frame(dfu_warpfield* wf, dfu_solver* solver, dfu_pointcache* canon_cache, const dfu_frame_params* p,
                         const uint16_t* depth_mm, size_t depth_pitch_bytes, uint16_t* dists, size_t dists_pitch_bytes,
                         const float* canon_v, unsigned long long canon_version, float* canon_warped, const float* live_v, int P,
                         dfu_stream stream) {
    DFU_REQUIRE(p && depth_mm && dists, DFU_ERR_INVALID, "NULL argument");
    DFU_REQUIRE(p->volume != nullptr, DFU_ERR_INVALID, "no volume");
    int rc = dfu_compute_dists(depth_mm, depth_pitch_bytes, dists, dists_pitch_bytes, p->rows, p->cols, p->intr, stream);
    if (rc != DFU_OK) return rc;
    if (wf && solver && P > 0) {
        DFU_REQUIRE(canon_cache && canon_v && canon_warped && live_v, DFU_ERR_INVALID, "the solve needs the canonical / live vertices");
        rc = dfu_warpfield_warp_cached(wf, canon_cache, canon_version, canon_v, nullptr, P, canon_warped, nullptr, p->blend_mode,
                                       DFU_NORMAL_REF, stream);
        if (rc != DFU_OK) return rc;
        rc = dfu_solver_init_problem(solver, canon_warped, nullptr, live_v, nullptr, P, nullptr, stream);
        if (rc != DFU_OK) return rc;
        rc = dfu_solver_solve_all(solver, stream);
        if (rc != DFU_OK) return rc;
    }
    return dfu_tsdf_integrate(p->volume, p->dims, p->voxel_size, p->trunc_dist, p->max_weight, p->vol2cam, p->intr, dists,
                              dists_pitch_bytes, p->rows, p->cols, wf, p->blend_mode, p->z0, p->z1, stream);
}
