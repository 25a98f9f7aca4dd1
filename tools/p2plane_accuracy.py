#!/usr/bin/env python
"""Distance of the point-to-plane SE(3) solve from the double-precision oracle on the parity-test scene, per execution path
(DFU_SOLVER_PATH) and recurrence-refresh period (DFU_P2P_REFRESH): how much margin the 1e-4 parity bound has."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dynfu_b200 as dfu  # noqa: E402
from oracle import pyoracle  # noqa: E402
from tools import synth  # noqa: E402
from tests.test_oracle_p2plane import rigid_scene  # noqa: E402


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).to("cuda", dtype=torch.float32)


def main():
    o = pyoracle.Oracle("brute")
    for n_nodes, n_pts, seed, tol in ((64, 6000, 7, 1e-7), (64, 6000, 7, 1e-9), (64, 6000, 7, 1e-11), (64, 6000, 3, 1e-9), (128, 12000, 5, 1e-9)):
        pos, dg_w, canon, live, live_n, R, t = rigid_scene(n_nodes=n_nodes, n_pts=n_pts, seed=seed)
        live = (live + np.random.default_rng(1).normal(0, 0.002, live.shape)).astype(np.float32)
        N = len(pos)
        prm_o = pyoracle.default_params(num_iter=4, nonlinear_iter=3, linear_iter=300, lambda_=200.0, psi_data=1.0, pcg_tol=1e-12)
        X_o, dq_o, st_o = o.solve_p2plane(pos, synth.identity_dq(N), dg_w, canon, live, live_n, prm_o)
        for path, refresh in (("multi", None), ("persistent", 16), ("persistent", 1)):
            os.environ["DFU_SOLVER_PATH"] = path
            os.environ.pop("DFU_P2P_REFRESH", None)
            if refresh:
                os.environ["DFU_P2P_REFRESH"] = str(refresh)
            wf = dfu.Warpfield("cuda:0")
            wf.init(0.08, dev(pos), dev(synth.identity_dq(N)), dev(dg_w))
            prm = dfu.CombinedSolverParameters(numIter=4, nonLinearIter=3, linearIter=300, earlyOut=False, pcgTolerance=tol)
            s = dfu.CombinedSolver(wf, prm, 4.652, 1.0, 200.0, 1e-4)
            s.setEnergy(s.ENERGY_P2PLANE_SE3)
            s.initializeProblemInstance(dev(canon), dev(live), liveNormals=dev(live_n))
            s.solveAll()
            st = s.getStats()
            X_g = s.getIncrements().cpu().numpy().astype(np.float64)
            print(json.dumps({"nodes": N, "points": n_pts, "seed": seed, "tol": tol, "path": path, "refresh": refresh,
                              "dX_max": float(np.max(np.abs(X_g - X_o))), "dE_rel": abs(st["final_energy"] - st_o[1]) / st_o[1],
                              "pcg_iterations": st["pcg_iterations"]}), flush=True)


if __name__ == "__main__":
    main()
