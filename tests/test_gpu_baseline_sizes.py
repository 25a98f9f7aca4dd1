"""GPU parity tests at the sizes BASELINE.json's configs name (run on the B200 box with -m gpu), against the CPU ORACLE
(kNN through the reference's own nanoflann when oracle/_ref is present):

    C1  deforming sphere, 640x480 depth, 256^3 volume, 1024 nodes: 5x10 solve + warped TSDF fusion
    C2  warped TSDF integration only, 512^3 volume, 4096 nodes
    C3  the full per-frame loop of bench.py (bending cylinder, 512^3, 4096 nodes, 75 852 points, 5 GN x 10 PCG)
    C5  data-term solve stress: 1280x720, ~300 k surface points, 32 768 nodes, 10 GN x 10 PCG
(C4, 1024^3 / 16 k nodes, is the multi-GPU configuration: its per-rank slab is covered by tests/mgpu_worker.py.)
Bars: packed TSDF voxels bit for bit; energies and node translations within 1e-4 relative (north_star)."""
import numpy as np
import pytest
import torch

from oracle import pyoracle
from tests import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dfu():
    import dynfu_b200

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return dynfu_b200


@pytest.fixture(scope="module")
def orc():
    """all host threads; nanoflann kNN when the reference-derived library is there"""
    import os

    pyoracle.build()
    try:
        o = pyoracle.Oracle("nanoflann")
    except FileNotFoundError:
        o = pyoracle.Oracle("brute")
    o.set_num_threads(len(os.sched_getaffinity(0)))
    return o


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to("cuda", dtype=dtype)


def _volume(dfu, dim):
    vol = dfu.TsdfVolume((dim, dim, dim))
    vol.setTruncDist(synth.TRUNC)
    vol.setMaxWeight(synth.MAX_WEIGHT)
    pose = np.eye(4)
    pose[:3, 3] = synth.VOLUME_T
    vol.setPose(pose)
    return vol


def _solve_both(dfu, orc, pos, dq, dg_w, eps, canon, live, gn, pcg, lambda_=200.0):
    prm_o = pyoracle.default_params(num_iter=gn, nonlinear_iter=1, linear_iter=pcg, lambda_=lambda_, pcg_tol=0.0, early_out=0)
    t_o, dq_o, st_o = orc.solve(pos, dq, dg_w, canon, live, prm_o)
    wf = dfu.Warpfield()
    wf.init(eps, dev(pos), dev(dq), dev(dg_w))
    prm = dfu.CombinedSolverParameters(numIter=gn, nonLinearIter=1, linearIter=pcg, earlyOut=False, pcgTolerance=0.0)
    s = dfu.CombinedSolver(wf, prm, 4.652, 1e-2, lambda_, 1e-4)
    s.initializeProblemInstance(dev(canon), dev(live))
    s.solveAll()
    st = s.getStats()
    t_g = s.getTranslations().cpu().numpy().astype(np.float64)
    assert st["gn_steps"] == gn and st["pcg_iterations"] == gn * pcg
    assert abs(st["initial_energy"] - st_o[0]) <= 1e-4 * st_o[0], (st, st_o)
    assert abs(st["final_energy"] - st_o[1]) <= 1e-4 * st_o[1], (st, st_o)
    scale = np.abs(t_o).max()
    assert np.max(np.abs(t_g - t_o)) <= 1e-4 * scale, np.max(np.abs(t_g - t_o)) / scale
    dq_g = wf.getNodes()[1].cpu().numpy()
    assert np.max(np.abs(dq_g - dq_o)) <= 1e-4 * scale
    return wf, dq_g, st, st_o


def _integrate_both(dfu, orc, dim, depth, wf, nodes, intr=synth.INTR, frames=1):
    vs = synth.voxel_size(dim)
    vol = _volume(dfu, dim)
    d = dfu.compute_dists(dev(depth.view(np.int16), torch.int16), intr)
    d_np = orc.compute_dists(depth, intr)
    assert np.array_equal(d.cpu().numpy().view(np.uint16), d_np)
    ref = np.zeros((dim,) * 3, np.uint32)
    for _ in range(frames):
        vol.integrate(d, np.eye(4), intr, wf)
        touched = orc.tsdf_integrate(ref, vs, vol.getTruncDist(), synth.MAX_WEIGHT, synth.VOL2CAM, intr, d_np, nodes=nodes)
        assert touched > 10000
    got = vol.data.cpu().numpy().view(np.uint32)
    bad = int(np.count_nonzero(got != ref))
    assert bad == 0, "%d of %d voxels differ from the oracle" % (bad, ref.size)


def test_c1_sphere_256_solve_and_warped_fusion(dfu, orc):
    """configs[0]: deforming sphere, 256^3, 1024 nodes -- solve (5 x 10) against the oracle, then warped fusion of the live
    depth through the solved field, packed voxels bit for bit"""
    eps = 0.025
    pos, _, dg_w, t_true = synth.sphere_nodes(1024, eps)
    depth = synth.sphere_depth()
    canon = synth.backproject(depth, synth.INTR)[::2]  # every 2nd valid pixel
    assert 20000 < len(canon) < 40000
    _, ties = orc.knn(pos, canon)
    assert ties == 0
    live = orc.warp(pos, synth.translations_to_dq(0.2 * t_true), dg_w, canon)
    wf, dq_g, _, _ = _solve_both(dfu, orc, pos, synth.identity_dq(1024), dg_w, eps, canon, live, 5, 10)
    _integrate_both(dfu, orc, 256, synth.sphere_depth(bump=0.004), wf, (pos, dq_g, dg_w), frames=2)


def test_c2_warped_integration_512_bit_exact(dfu, orc):
    """configs[1]: warped TSDF integration only, 512^3, 4096 nodes, every voxel against the oracle"""
    pos, dq, dg_w, _ = synth.sphere_nodes(4096, 0.0125)
    wf = dfu.Warpfield()
    wf.init(0.0125, dev(pos), dev(dq), dev(dg_w))
    _integrate_both(dfu, orc, 512, synth.sphere_depth(), wf, (pos, dq, dg_w))


def test_c3_bench_frame_solve_and_fusion(dfu, orc):
    """configs[2] = bench.py's frame: bending cylinder, 4096 nodes, 75 852 surface points, 5 GN x 10 PCG, lambda 200 --
    energies and translations within 1e-4 of the oracle, then the warped fusion of the bent depth at 512^3 bit for bit"""
    import bench

    sc = bench.make_scene()
    assert len(sc["canon"]) == 75852 and len(sc["pos"]) == 4096
    wf, dq_g, st, st_o = _solve_both(dfu, orc, sc["pos"], sc["dq"], sc["dg_w"], bench.EPSILON, sc["canon"], sc["lives"][0],
                                     bench.GN_ITERS, bench.PCG_ITERS, bench.LAMBDA)
    assert st["final_energy"] < 0.05 * st["initial_energy"]
    _integrate_both(dfu, orc, 512, sc["depths"][0], wf, (sc["pos"], dq_g, sc["dg_w"]))


def test_c5_solve_stress_against_the_oracle(dfu, orc):
    """configs[4]: 1280x720, ~300 k points, 32 768 nodes, 10 GN x 10 PCG -- the GPU solve (the generic explicit-matrix path:
    more rows than the register version holds) against the ORACLE, both energies and every translation"""
    eps = 0.004
    pos, dq, dg_w = synth.cylinder_nodes(256, 128, eps)
    # 300 000 surface points sampled uniformly on the camera-facing half of the cylinder, volume-local metres
    rng = np.random.default_rng(synth.SEED + 8)  # a seed without bit-equal kNN distances (asserted below)
    th = np.pi + rng.uniform(0.0, np.pi, 300000)
    c_vol = np.array([0.0, 0.0, 2.0]) - synth.VOLUME_T
    canon = np.stack([c_vol[0] + 0.3 * np.cos(th), c_vol[1] + rng.uniform(-0.8, 0.8, 300000), c_vol[2] + 0.3 * np.sin(th)],
                     -1).astype(np.float32)
    live = synth.bend(canon, 0.005 * eps / 0.0125)
    _, ties = orc.knn(pos, canon)
    assert len(pos) == 32768 and len(canon) == 300000 and ties == 0
    _solve_both(dfu, orc, pos, dq, dg_w, eps, canon, live, 10, 10)
