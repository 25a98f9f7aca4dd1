"""CPU tests of the oracle's kNN / blend / warp / TSDF restatement: pins kNN to the reference's own
nanoflann (oracle/_ref) and checks the size-independent properties the GPU tests reuse."""
import numpy as np
import pytest

from oracle import pyoracle
from tests import synth


def test_knn_bruteforce_equals_reference_nanoflann(oracle, oracle_nf):
    """(dist, idx) brute force == the reference's KD-tree (include/nanoflann/nanoflann.hpp) whenever no
    two of a query's 9 nearest distances are bit-equal."""
    rng = np.random.default_rng(synth.SEED)
    nodes, _, _, _ = synth.sphere_nodes(4096, 0.0125)
    q = (nodes[rng.integers(0, 4096, 20000)] + rng.normal(0, 0.05, (20000, 3))).astype(np.float32)
    idx_b, d_b, ties = oracle.knn(nodes, q, return_dist=True)
    assert ties == 0
    idx_n, d_n, _ = oracle_nf.knn(nodes, q, return_dist=True)
    assert np.array_equal(idx_b, idx_n)
    assert np.array_equal(d_b, d_n)  # bit-exact squared distances
    assert np.all(np.diff(d_b, axis=1) >= 0)


def test_knn_reference_fixture_nodes(oracle, oracle_nf):
    from tests import fixtures_opt as fx
    q = np.array([(0, 0.04, 0), (2, 2, 2), (10.5, 10.5, 10.5)], np.float32)
    ib, _ = oracle.knn(fx.ALL_NODES, q)
    inn, _ = oracle_nf.knn(fx.ALL_NODES, q)
    # integer lattice nodes DO tie; only tie-free prefixes are comparable with the KD-tree's visitation order
    _, d, _ = oracle.knn(fx.ALL_NODES, q, return_dist=True)
    for r in range(3):
        strict = np.concatenate([[True], np.diff(d[r]) > 0])
        n_ok = int(np.argmin(strict)) if not strict.all() else 8
        assert np.array_equal(ib[r, :max(n_ok - 1, 0)], inn[r, :max(n_ok - 1, 0)])


def test_blend_identity_rotation_is_weighted_translation(oracle):
    """with identity rotations REF_COMPOSE degenerates to v' = v + sum_k w_k t_k."""
    nodes, dq, dg_w, t = synth.sphere_nodes(1024, 0.025)
    rng = np.random.default_rng(3)
    pts = (nodes[rng.integers(0, 1024, 500)] + rng.normal(0, 0.02, (500, 3))).astype(np.float32)
    warped = oracle.warp(nodes, dq, dg_w, pts)
    idx, _ = oracle.knn(nodes, pts)
    w = np.array([[oracle.node_weight(nodes[j], dg_w[j], pts[v]) for j in idx[v]] for v in range(500)], np.float64)
    expect = pts + np.einsum("vk,vkc->vc", w, t[idx].astype(np.float64))
    assert np.max(np.abs(warped - expect)) < 2e-6


def test_dqb_sum_rigid_field(oracle):
    """True DQB of identical node transforms is that transform (north-star mode)."""
    nodes, _, dg_w, _ = synth.sphere_nodes(256, 0.05)
    one = oracle.dq_from_euler(0.3, -0.2, 0.1, 0.05, -0.02, 0.03)
    dq = np.tile(one, (256, 1))
    pts = nodes[:50] + 0.01
    got = oracle.warp(nodes, dq, dg_w, pts, blend_mode=pyoracle.BLEND_DQB_SUM)
    exp = np.array([oracle.dq_transform_vertex(one, p) for p in pts])
    assert np.max(np.abs(got - exp)) < 1e-5
    # far from every node the support underflows to 0 -> identity
    far = np.array([[100.0, 100.0, 100.0]], np.float32)
    assert np.array_equal(oracle.warp(nodes, dq, dg_w, far, blend_mode=pyoracle.BLEND_DQB_SUM), far)


def test_normals_ref_mode_adds_translation(oracle):
    """dual_quaternion.hpp:217-228 quirk: normals get the vertex formula, translation included."""
    dq = oracle.dq_from_euler(0, 0, 0, 1, 2, 3)
    assert np.allclose(oracle.dq_transform_normal(dq, [0, 0, 1]), [1, 2, 4])
    assert np.allclose(oracle.dq_transform_normal(dq, [0, 0, 1], pyoracle.NORMAL_ROTATE_ONLY), [0, 0, 1])


def test_compute_dists_matches_formula(oracle):
    depth = synth.sphere_depth()
    d = oracle.compute_dists(depth, synth.INTR)
    y, x = 240, 320
    lam = np.sqrt(np.float32(((x - 319.5) / 525.0) ** 2 + ((y - 239.5) / 525.0) ** 2 + 1))
    assert abs(oracle.half2float(int(d[y, x])) - depth[y, x] * lam * 1e-3) < 2e-3
    assert d[0, 0] == 0 and depth[0, 0] == 0


def _vol(dim):
    return np.zeros((dim, dim, dim), np.uint32)


def _integrate(o, vol, dists, nodes=None, **kw):
    dim = vol.shape[0]
    vs = synth.voxel_size(dim)
    trunc = o.trunc_dist(synth.TRUNC, vs)
    return o.tsdf_integrate(vol, vs, trunc, synth.MAX_WEIGHT, synth.VOL2CAM, synth.INTR, dists, nodes=nodes, **kw)


@pytest.fixture(scope="module")
def dists(oracle):
    return oracle.compute_dists(synth.sphere_depth(), synth.INTR)


def test_tsdf_rigid_sanity(oracle, dists):
    vol = _vol(64)
    touched = _integrate(oracle, vol, dists)
    assert touched > 1000
    tsdf = (vol & 0xffff).astype(np.uint16).view(np.float16).astype(np.float32)
    w = vol >> 16
    assert set(np.unique(w)) <= {0, 1}
    assert tsdf.max() <= 1.0 and tsdf.min() >= -1.0
    # voxel in front of the sphere on the optical axis is free space (tsdf == 1), behind the surface untouched
    vs = 3.0 / 64
    zi_front = int((1.2 - 0.5) / vs)
    assert tsdf[zi_front, 32, 32] == 1.0 and w[zi_front, 32, 32] == 1
    zi_in = int((2.0 - 0.5) / vs)
    assert w[zi_in, 32, 32] == 0


def test_tsdf_identity_warp_equals_rigid(oracle, dists):
    """with identity node transforms the warped integrator IS the rigid one."""
    nodes, _, dg_w, _ = synth.sphere_nodes(256, 0.05)
    a, b = _vol(48), _vol(48)
    _integrate(oracle, a, dists)
    _integrate(oracle, b, dists, nodes=(nodes, synth.identity_dq(256), dg_w))
    assert np.array_equal(a, b)
    c = _vol(48)
    _integrate(oracle, c, dists, nodes=(nodes, synth.identity_dq(256), dg_w), blend_mode=pyoracle.BLEND_DQB_SUM)
    assert np.array_equal(a, c)


def test_tsdf_slab_invariance_and_running_average(oracle, dists):
    nodes, dq, dg_w, _ = synth.sphere_nodes(256, 0.05)
    full, slabs = _vol(48), _vol(48)
    _integrate(oracle, full, dists, nodes=(nodes, dq, dg_w))
    for z0 in range(0, 48, 12):
        _integrate(oracle, slabs, dists, nodes=(nodes, dq, dg_w), z0=z0, z1=z0 + 12)
    assert np.array_equal(full, slabs)
    rigid = _vol(48)
    _integrate(oracle, rigid, dists)
    assert not np.array_equal(full, rigid)  # the warp does something
    # second frame: weights go to 2 where touched twice, clamp at max_weight
    _integrate(oracle, full, dists, nodes=(nodes, dq, dg_w))
    assert (full >> 16).max() == 2
    oracle.tsdf_clear(full)
    assert not full.any()


def test_trunc_dist_clamp(oracle):  # tsdf_volume.cpp:57-61
    assert oracle.trunc_dist(0.04, synth.voxel_size(512)) == np.float32(0.04)
    vs = synth.voxel_size(64)
    assert oracle.trunc_dist(0.04, vs) == np.float32(2.1) * vs[0]
