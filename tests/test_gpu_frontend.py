"""GPU parity tests for the rows either side of the hot path: depth -> points/normals (f3),
valid-pixel compaction, live<->canonical correspondences (f2).  Bar: bit-exact (these are index / IEEE-float
pipelines); everything goes through the C-ABI."""
import numpy as np
import pytest
import torch

from tests import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fe():
    import dynfu_b200
    from dynfu_b200 import frontend

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return frontend


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to("cuda", dtype=dtype)


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def holey_depth(rows, cols, seed=3):
    rng = np.random.default_rng(seed)
    intr = synth.intr_for(cols, rows)
    depth = synth.sphere_depth(rows, cols, intr, bump=0.02).copy()
    depth[rng.random((rows, cols)) < 0.02] = 0  # sensor drop-outs
    depth[: rows // 8, : cols // 3] = 1234       # a flat background patch
    return depth, intr


# ------------------------------------------------------------------------------------------------ f3
@pytest.mark.parametrize("rows,cols", [(480, 640), (61, 47), (2, 2), (1, 5)])
def test_points_normals_bit_exact(fe, oracle, rows, cols):
    depth, intr = holey_depth(rows, cols)
    p_o, n_o = oracle.points_normals(depth, intr)
    p_g, n_g = fe.compute_points_normals(dev(depth.view(np.int16), torch.int16), intr)
    p_g, n_g = p_g.cpu().numpy(), n_g.cpu().numpy()
    valid = ~np.isnan(p_o[..., 0])
    assert np.array_equal(valid, ~np.isnan(p_g[..., 0]))
    assert np.array_equal(valid, ~np.isnan(n_g[..., 0]))
    assert np.isnan(p_g[~valid]).all() and np.isnan(n_g[~valid]).all()
    assert same_bits(p_g[valid], p_o[valid])
    assert same_bits(n_g[valid], n_o[valid])
    if valid.any():  # sanity of the oracle itself: unit normals facing the camera
        assert np.allclose(np.linalg.norm(n_o[valid][:, :3], axis=1), 1.0, atol=1e-5)


def test_points_normals_pitched_views(fe, oracle):
    """row pitch != cols * elem size on every image, like DeviceArray2D (include/kfusion/cuda/device_array.hpp)"""
    depth, intr = holey_depth(120, 160)
    big = torch.zeros((120, 192), dtype=torch.int16, device="cuda")
    big[:, :160] = dev(depth.view(np.int16), torch.int16)
    pts = torch.zeros((120, 200, 4), device="cuda")[:, :160]
    nrm = torch.zeros((120, 176, 4), device="cuda")[:, :160]
    fe.compute_points_normals(big[:, :160], intr, pts, nrm)
    p_o, n_o = oracle.points_normals(depth, intr)
    assert np.array_equal(np.nan_to_num(pts.cpu().numpy(), nan=7.0), np.nan_to_num(p_o, nan=7.0))
    assert np.array_equal(np.nan_to_num(nrm.cpu().numpy(), nan=7.0), np.nan_to_num(n_o, nan=7.0))


@pytest.mark.parametrize("with_xform", [False, True])
def test_compact_points_raster_order(fe, oracle, with_xform):
    depth, intr = holey_depth(480, 640)
    p_o, n_o = oracle.points_normals(depth, intr)
    xf = None
    if with_xform:  # camera -> volume-local: inverse of the default volume pose, plus a small rotation
        a = 0.1
        xf = np.eye(4)
        xf[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
        xf[:3, 3] = [1.5, 1.5, -0.5]
    xf12 = None if xf is None else np.concatenate([xf[:3, :3].reshape(-1), xf[:3, 3]]).astype(np.float32)
    v_o, m_o = oracle.compact_points(p_o, n_o, xf12)
    p_g, n_g = fe.compute_points_normals(dev(depth.view(np.int16), torch.int16), intr)
    v_g, m_g = fe.compact_points(p_g, n_g, xf)
    assert v_g.shape[0] == v_o.shape[0] > 1000
    assert same_bits(v_g.cpu().numpy(), v_o)
    assert same_bits(m_g.cpu().numpy(), m_o)
    # points only, device-side count, truncated capacity
    v2, _, cnt = fe.compact_points(p_g, None, xf, capacity=500, sync=False)
    assert int(cnt.item()) == v_o.shape[0]
    assert same_bits(v2.cpu().numpy(), v_o[:500])


def test_compact_points_empty(fe):
    pts = torch.full((16, 16, 4), float("nan"), device="cuda")
    v, n = fe.compact_points(pts, pts.clone())
    assert v.shape[0] == 0 and n.shape[0] == 0


# ------------------------------------------------------------------------------------------------ f2
def surface_points(n, seed, noise=0.0):
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p = np.array([1.5, 1.5, 1.5]) + 0.5 * d + rng.normal(0, noise, (n, 3)) if noise else np.array([1.5, 1.5, 1.5]) + 0.5 * d
    return p.astype(np.float32)


@pytest.mark.parametrize("kind,P,Q", [("surface", 76800, 76800), ("surface", 3000, 10000), ("cube", 20000, 5000),
                                      ("line", 5000, 2000), ("surface", 1, 100), ("surface", 7, 3)])
def test_nearest_bit_exact(fe, oracle, kind, P, Q):
    rng = np.random.default_rng(P + Q)
    if kind == "surface":
        pts = surface_points(P, 1)
        q = (surface_points(Q, 2) * 1.0 + rng.normal(0, 0.01, (Q, 3))).astype(np.float32)
    elif kind == "cube":
        pts = rng.uniform(0, 3, (P, 3)).astype(np.float32)
        q = rng.uniform(-0.5, 3.5, (Q, 3)).astype(np.float32)
    else:  # degenerate extent in two axes
        pts = np.zeros((P, 3), np.float32)
        pts[:, 0] = rng.uniform(0, 3, P)
        q = rng.uniform(-1, 4, (Q, 3)).astype(np.float32)
    idx_o, d_o, ties = oracle.knn(pts, q, k=1, return_dist=True)
    index = fe.PointIndex().build(dev(pts))
    idx_g, d_g = index.nearest(dev(q), return_dist=True)
    assert ties == 0 or kind == "line"  # collinear points do produce equal distances; key (dist2, idx) settles them
    assert np.array_equal(idx_g.cpu().numpy(), idx_o[:, 0])
    assert same_bits(d_g.cpu().numpy(), d_o[:, 0])


def test_nearest_far_queries_and_ties(fe, oracle):
    """queries far outside the indexed cloud (the sweep path) and exact duplicates (lower index wins)"""
    pts = surface_points(4000, 5)
    pts = np.concatenate([pts, pts[:500]])  # 500 duplicated points, indices 4000.. duplicate 0..499
    rng = np.random.default_rng(9)
    q = np.concatenate([pts[:600], rng.uniform(-20, 20, (3000, 3)).astype(np.float32)])
    idx_o, d_o, _ = oracle.knn(pts, q, k=1, return_dist=True)  # brute force, key (dist2, idx)
    index = fe.PointIndex().build(dev(pts))
    idx_g, d_g = index.nearest(dev(q), return_dist=True)
    assert np.array_equal(idx_g.cpu().numpy(), idx_o[:, 0])
    assert same_bits(d_g.cpu().numpy(), d_o[:, 0])
    assert (idx_g.cpu().numpy()[:500] == np.arange(500)).all()


def test_find_corresponding_matches_reference_kdtree(fe, oracle_nf):
    """DynFusion::findCorrespondingFrame against the reference's own nanoflann (oracle/_ref)"""
    rng = np.random.default_rng(11)
    canon = surface_points(30000, 21)
    normals = ((canon - 1.5) / 0.5).astype(np.float32)
    live = (surface_points(25000, 22) + rng.normal(0, 0.004, (25000, 3))).astype(np.float32)
    v_o, n_o, idx_o, _ = oracle_nf.find_corresponding(canon, normals, live)
    index = fe.PointIndex()
    v_g, n_g, idx_g = index.find_corresponding(dev(canon), dev(normals), dev(live), return_index=True)
    assert np.array_equal(idx_g.cpu().numpy(), idx_o)
    assert same_bits(v_g.cpu().numpy(), v_o)
    assert same_bits(n_g.cpu().numpy(), n_o)
    # the handle is reusable with a different cloud size (per-frame rebuild)
    v2, _ = index.find_corresponding(dev(canon[:1000]), None, dev(live[:50]))
    v2_o, _, _, _ = oracle_nf.find_corresponding(canon[:1000], None, live[:50])
    assert same_bits(v2.cpu().numpy(), v2_o)


def test_depth_to_solver_pipeline(fe, oracle):
    """depth -> points -> volume-local compaction -> correspondences -> shapes the solver accepts"""
    depth, intr = holey_depth(240, 320)
    p, n = fe.compute_points_normals(dev(depth.view(np.int16), torch.int16), intr)
    cam2vol = np.eye(4)
    cam2vol[:3, 3] = [1.5, 1.5, -0.5]
    live_v, live_n = fe.compact_points(p, n, cam2vol)
    canon = live_v[::2].contiguous() + 0.002
    cv, cn = fe.find_corresponding(canon, live_n[::2].contiguous(), live_v)
    assert cv.shape == live_v.shape and cn.shape == live_v.shape
    # every live vertex is within ~2 pixels' footprint of its partner
    assert float((cv - live_v).norm(dim=1).max()) < 0.05
