#!/usr/bin/env python
"""Device time of the point-to-plane SE(3) solve (north-star extension) at the headline size (4096 nodes, ~76 k points,
5 GN x 10 PCG) for each execution path: one kernel per phase, and the one-launch cooperative kernel at several grid sizes.
One JSON line per setting; with DFU_SOLVER_PROFILE=1 the library prints the per-phase cycles of CTA 0 to stderr."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dynfu_b200 as dfu  # noqa: E402
from tools.bench_configs import DEV, dev, scene, timed  # noqa: E402


def main():
    eps, gn, pcg = 0.0125, 5, 10
    sc = scene(512, 64, 64, eps)
    nn = sc["canon"].astype(np.float64) - np.array([1.5, 0.0, 1.5])
    nn[:, 1] = 0.0
    nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    live_n = dev(nn.astype(np.float32))
    canon, live = dev(sc["canon"]), dev(sc["lives"][0])
    dq0 = dev(sc["dq"])
    settings = [("multi", None), ("persistent", None), ("persistent", 148), ("persistent", 74)]
    for path, ctas in settings:
        os.environ["DFU_SOLVER_PATH"] = path
        os.environ.pop("DFU_P2P_CTAS", None)
        if ctas:
            os.environ["DFU_P2P_CTAS"] = str(ctas)
        wf = dfu.Warpfield(DEV)
        wf.init(eps, dev(sc["pos"]), dev(sc["dq"]), dev(sc["w"]))
        prm = dfu.CombinedSolverParameters(numIter=gn, nonLinearIter=1, linearIter=pcg, earlyOut=False, pcgTolerance=0.0)
        s = dfu.CombinedSolver(wf, prm, 4.652, 1e-2, 200.0, 1e-4)
        s.setEnergy(s.ENERGY_P2PLANE_SE3)

        def solve(i):
            if i % 4 == 0:  # the increments are composed onto the nodes by every solve: restart before they pile up
                wf.setTransformations(dq0)
            s.solveAll()

        s.initializeProblemInstance(canon, live, liveNormals=live_n)
        ms = timed(solve, n=40, warm=4)
        st = s.getStats()
        print(json.dumps({"path": path, "ctas": ctas or "default", "solve_ms": round(ms, 4), "points": int(canon.shape[0]),
                          "nodes": int(sc["pos"].shape[0]), "gn_x_pcg": "%dx%d" % (gn, pcg),
                          "solver": {k: st[k] for k in ("initial_energy", "final_energy", "pcg_iterations", "gn_steps")}}), flush=True)
        if path == "persistent" and os.environ.get("DFU_P2P_PROFILE_ONCE", "1") == "1":
            os.environ["DFU_SOLVER_PROFILE"] = "1"
            s.solveAll()
            torch.cuda.synchronize()
            os.environ.pop("DFU_SOLVER_PROFILE")
        del s, wf


if __name__ == "__main__":
    main()
