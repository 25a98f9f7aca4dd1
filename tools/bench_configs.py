#!/usr/bin/env python
"""Device times of the BASELINE.json configurations other than the headline one (which bench.py measures), on ONE GPU:
   C1  256^3, ~1k nodes           full frame (kNN + DQB + 5x10 solve + warped integrate)
   C2  512^3, 4096 nodes          warped integration only
   C4  1024^3, 16k nodes          the per-GPU share at 8 GPUs: a 128-plane z-slab (+ the replicated solve)
   C5  1280x720, ~300k points, 32k nodes   data-term solve stress, 10 GN x 10 PCG
Prints one JSON line per configuration; the numbers go into BASELINE.md.  CUDA-event timing, 3 warm-up frames."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dynfu_b200 as dfu  # noqa: E402
from tools import synth  # noqa: E402

DEV = torch.device("cuda", 0)


def dev(a, dt=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to(DEV, dtype=dt)


def scene(dim, n_theta, n_y, eps, rows=480, cols=640):
    # bend amplitude in proportion to the node radius: the reference's frame loop only tracks deformations well below dg_w
    # (DESIGN.md section 2); bench.py's scene has kappa = 0.005 at epsilon = 0.0125
    kappa = 0.005 * eps / 0.0125
    intr = synth.intr_for(cols, rows)
    depth0 = synth.cylinder_depth(rows, cols, intr)
    canon = synth.backproject(depth0, intr)
    pos, dq, w = synth.cylinder_nodes(n_theta, n_y, eps)
    depths = [synth.cylinder_depth(rows, cols, intr, kappa=kappa * a) for a in (1.0, 2.0)]
    lives = [synth.bend(canon, kappa * a) for a in (1.0, 2.0)]
    return dict(intr=intr, depth0=depth0, canon=canon, pos=pos, dq=dq, w=w, depths=depths, lives=lives, rows=rows, cols=cols)


def timed(fn, n=20, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(warm + i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def run(name, dim, n_theta, n_y, eps, gn, pcg, rows=480, cols=640, z0=0, z1=None, do_solve=True, do_integrate=True, p2plane=False):
    sc = scene(dim, n_theta, n_y, eps, rows, cols)
    kp = dfu.KinFuParams(cols=cols, rows=rows, intr=tuple(float(x) for x in sc["intr"]), volume_dims=(dim, dim, dim))
    prm = dfu.DynFuParams(kinfuParams=kp, epsilon=eps, lambda_=200.0,
                          blend_mode=dfu.BLEND_DQB_SUM if p2plane else dfu.BLEND_REF_COMPOSE,  # rotations: true DQ blending
                          solver=dfu.CombinedSolverParameters(numIter=gn, nonLinearIter=1, linearIter=pcg, earlyOut=False,
                                                              pcgTolerance=0.0))
    df = dfu.DynFusion(prm, device=DEV, z0=z0, z1=z1)
    df.init(dev(sc["canon"]), None, nodes=(dev(sc["pos"]), dev(sc["dq"]), dev(sc["w"])))
    depth_dev = [torch.from_numpy(d.view(np.int16)).to(DEV) for d in sc["depths"]]
    live_dev = [dev(l) for l in sc["lives"]]
    df(torch.from_numpy(sc["depth0"].view(np.int16)).pin_memory())
    t = {}
    live_n = None
    if p2plane:  # north-star extension: point-to-plane SE(3) data term; normals of the (bent) cylinder ~ radial
        c = np.array([1.5, 1.5, 1.5]) + np.array([0.0, 0.0, 0.0])
        nn = sc["canon"].astype(np.float64) - np.array([1.5, 0.0, 1.5])
        nn[:, 1] = 0.0
        nn /= np.linalg.norm(nn, axis=1, keepdims=True)
        live_n = dev(nn.astype(np.float32))
        df.solver.setEnergy(df.solver.ENERGY_P2PLANE_SE3)

    def solve(i):
        df.canonicalWarpedToLive, _ = df.warpCanonical()
        df.solver.initializeProblemInstance(df.canonicalWarpedToLive, live_dev[i % 2], liveNormals=live_n)
        df.solver.solveAll()

    def integrate(i):
        dfu.compute_dists(depth_dev[i % 2], kp.intr, out=df._dists)
        df.volume.integrate(df._dists, df.camera_pose, kp.intr, df.warpfield, prm.blend_mode)

    def frame(i):
        if do_solve:
            solve(i)
        if do_integrate:
            integrate(i)

    t["frame_ms"] = timed(frame)
    if do_solve:
        t["points_solve_ms"] = timed(solve)
    if do_integrate:
        t["integrate_ms"] = timed(integrate)
    planes = (z1 if z1 is not None else dim) - z0
    vox = dim * dim * planes
    out = {"config": name, "volume": "%d^3" % dim, "planes": planes, "nodes": int(sc["pos"].shape[0]), "points": int(sc["canon"].shape[0]),
           "gn_x_pcg": "%dx%d" % (gn, pcg), "frames_per_s": 1e3 / t["frame_ms"], **{k: round(v, 4) for k, v in t.items()}}
    if do_integrate:
        out["voxels_per_s"] = vox / (t["integrate_ms"] * 1e-3)
        out["integrate_algorithmic_GBps"] = vox * 8 / (t["integrate_ms"] * 1e-3) / 1e9
    if do_solve:
        st = df.solver.getStats()
        out["solver"] = {k: st[k] for k in ("initial_energy", "final_energy", "pcg_iterations", "gn_steps")}
    pool, built = df.warpfield.cacheStats()
    out["voxel_cache_bricks"] = [pool, built]
    print(json.dumps(out), flush=True)
    del df
    torch.cuda.empty_cache()


if __name__ == "__main__":
    which = sys.argv[1:] or ["C1", "C2", "C4", "C5"]
    if "C1" in which:
        run("C1 full frame", 256, 32, 32, 0.025, 5, 10)
    if "C2" in which:
        run("C2 warped integration only", 512, 64, 64, 0.0125, 5, 10, do_solve=False)
    if "C3p" in which:  # the headline frame with the point-to-plane SE(3) term instead of the reference's energy
        run("C3 with the point-to-plane SE(3) data term (north-star extension)", 512, 64, 64, 0.0125, 5, 10, p2plane=True)
    if "C4" in which:  # the slab holding the surface's z range is the heaviest one: planes 384..512 of 1024 (z 1.125..1.5 m)
        run("C4 per-GPU share at 8 GPUs (128-plane slab + replicated solve)", 1024, 128, 128, 0.00625, 5, 10, z0=384, z1=512)
    if "C5" in which:
        run("C5 data-term solve stress", 256, 256, 128, 0.004, 10, 10, rows=720, cols=1280, do_integrate=False)
