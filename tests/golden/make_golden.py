#!/usr/bin/env python
"""Generates tests/golden/hotpath_small.npz -- committed golden input/output vectors of the hot path.

Run in the build container (where /root/reference exists): kNN goes through the reference's OWN nanoflann
(oracle/_ref), everything else through the oracle restatement that the reference's golden tests pin.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from tests import synth  # noqa: E402


def main():
    pyoracle.build()
    o = pyoracle.Oracle("nanoflann")  # reference KD-tree
    ob = pyoracle.Oracle("brute")
    rng = np.random.default_rng(20261017)
    pos, dq_t, dg_w, t_true = synth.sphere_nodes(256, 0.04)
    _, dq_r, _, _ = synth.sphere_nodes(256, 0.04, rotations=True)
    pts = (pos[rng.integers(0, 256, 400)] + rng.normal(0, 0.04, (400, 3))).astype(np.float32)
    nrm = rng.normal(size=(400, 3)).astype(np.float32)
    idx, d2, _ = o.knn(pos, pts, return_dist=True)
    idx_b, d2_b, ties = ob.knn(pos, pts, return_dist=True)
    assert ties == 0 and np.array_equal(idx, idx_b) and np.array_equal(d2, d2_b)
    out = dict(pos=pos, dq_t=dq_t, dq_r=dq_r, dg_w=dg_w, pts=pts, nrm=nrm, knn_idx=idx, knn_d2=d2)
    for name, dq in (("t", dq_t), ("r", dq_r)):
        for mode in (0, 1):
            out["blend_%s_%d" % (name, mode)] = o.blend(pos, dq, dg_w, pts, mode)
            v, n = o.warp(pos, dq, dg_w, pts, nrm, mode, 0)
            out["warp_v_%s_%d" % (name, mode)] = v
            out["warp_n_%s_%d" % (name, mode)] = n
    # TSDF: 32^3 volume, 160x120 depth, two frames, rigid / translation-only warp / rotations (compose) / DQB
    rows, cols = 120, 160
    intr = synth.intr_for(cols, rows)
    depth = synth.sphere_depth(rows, cols, intr)
    dists = o.compute_dists(depth, intr)
    out.update(depth=depth, dists=dists, intr=intr)
    vs = synth.voxel_size(32)
    tr = o.trunc_dist(synth.TRUNC, vs)
    for name, nodes, mode in (("rigid", None, 0), ("t0", (pos, dq_t, dg_w), 0), ("r0", (pos, dq_r, dg_w), 0),
                              ("r1", (pos, dq_r, dg_w), 1)):
        vol = np.zeros((32, 32, 32), np.uint32)
        for _ in range(2):
            o.tsdf_integrate(vol, vs, tr, synth.MAX_WEIGHT, synth.VOL2CAM, intr, dists, nodes=nodes, blend_mode=mode)
        out["tsdf_" + name] = vol
    # solver: well-posed problem, fixed 5 x 10 iterations and converged
    canon = (pos[rng.integers(0, 256, 3000)] + rng.normal(0, 0.01, (3000, 3))).astype(np.float32)
    live = o.warp(pos, synth.translations_to_dq(0.2 * t_true), dg_w, canon)
    out.update(canon=canon, live=live)
    for tag, prm in (("fixed", pyoracle.default_params(num_iter=5, nonlinear_iter=1, linear_iter=10, lambda_=200.0,
                                                       pcg_tol=0.0, early_out=0)),
                     ("conv", pyoracle.default_params(num_iter=6, nonlinear_iter=2, linear_iter=400, lambda_=200.0,
                                                      pcg_tol=1e-10))):
        t, dq_new, st = o.solve(pos, synth.identity_dq(256), dg_w, canon, live, prm)
        out["solve_t_" + tag] = t
        out["solve_stats_" + tag] = st
    path = os.path.join(ROOT, "tests", "golden", "hotpath_small.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
