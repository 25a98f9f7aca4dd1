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


# ------------------------------------------------------------------------------------------------ f1
def update_scene(n_nodes=300, rows=240, cols=320):
    """a sphere surface seen by the camera, sparsely covered by nodes -> part of it is unsupported"""
    intr = synth.intr_for(cols, rows)
    depth = synth.sphere_depth(rows, cols, intr, bump=0.01)
    pos, dq, w, _ = synth.sphere_nodes(n_nodes, 0.0125, rotations=True)
    return depth, intr, pos, dq, w


def make_wf(pos, dq, w, eps=0.0125):
    import dynfu_b200

    wf = dynfu_b200.Warpfield()
    wf.init(eps, dev(pos), dev(dq), dev(w))
    return wf


def frame_vertices(oracle, depth, intr):
    p, _ = oracle.points_normals(depth, intr)
    return oracle.compact_points(p, None, np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 1.5, 1.5, -0.5], np.float32))


def test_unsupported_vertices_bit_exact(fe, oracle):
    depth, intr, pos, dq, w = update_scene()
    v = frame_vertices(oracle, depth, intr)
    mask_o = oracle.unsupported(pos, w, v)
    assert 0 < mask_o.sum() < v.shape[0]  # the scene exercises both outcomes
    wf = make_wf(pos, dq, w)
    mask_g = wf.getUnsupportedVertices(dev(v), return_mask=True).cpu().numpy()
    assert np.array_equal(mask_g, mask_o)
    assert same_bits(wf.getUnsupportedVertices(dev(v)).cpu().numpy(), v[mask_o])
    # vertices exactly one radius away from their nearest node: ratio == 1 counts as unsupported (>=)
    edge = pos[:50] + np.float32([0.0375, 0, 0])
    assert np.array_equal(wf.getUnsupportedVertices(dev(edge), return_mask=True).cpu().numpy(), oracle.unsupported(pos, w, edge))


@pytest.mark.parametrize("kind", ["surface", "cube", "single", "negative"])
def test_voxel_grid_bit_exact(fe, oracle, kind):
    rng = np.random.default_rng(17)
    if kind == "surface":
        pts = surface_points(40000, 3)
    elif kind == "cube":
        pts = rng.uniform(0.2, 2.9, (30000, 3)).astype(np.float32)
    elif kind == "single":
        pts = np.float32([[0.31, 0.32, 0.33]])
    else:  # coordinates straddling zero: floor() towards -inf, negative min_b
        pts = rng.uniform(-1.0, 1.0, (20000, 3)).astype(np.float32)
    c_o = oracle.voxel_grid(pts, 0.05)
    c_g = fe.voxel_grid_filter(dev(pts), 0.05).cpu().numpy()
    assert c_g.shape == c_o.shape
    assert same_bits(c_g, c_o)


def test_voxel_grid_many_points_per_cell_and_limits(fe, oracle):
    rng = np.random.default_rng(23)
    pts = (np.float32([1.0, 1.0, 1.0]) + rng.uniform(0, 0.1, (20000, 3))).astype(np.float32)  # ~2500 points per cell
    c_o = oracle.voxel_grid(pts, 0.05)
    c_g = fe.voxel_grid_filter(dev(pts), 0.05).cpu().numpy()
    assert same_bits(c_g, c_o)
    # the unstable std::sort of PCL changes the float sums by a few ulp only (measured, not asserted bit-exact)
    assert np.abs(oracle.voxel_grid(pts, 0.05, order_mode=1) - c_o).max() < 1e-5
    import dynfu_b200
    with pytest.raises(dynfu_b200.DfuError):  # 2^21-cell table exceeded
        fe.voxel_grid_filter(dev(np.float32([[0, 0, 0], [100, 100, 100]])), 0.05)
    assert fe.voxel_grid_filter(dev(np.zeros((0, 3), np.float32)), 0.05).shape[0] == 0


@pytest.mark.parametrize("n_nodes", [300, 8])
def test_warpfield_update_bit_exact(fe, oracle, n_nodes):
    depth, intr, pos, dq, w = update_scene(n_nodes)
    v = frame_vertices(oracle, depth, intr)
    po, qo, wo = oracle.warpfield_update(pos, dq, w, 0.0125, v)
    wf = make_wf(pos, dq, w)
    nu, nn = wf.update(dev(v))
    assert nu == int(oracle.unsupported(pos, w, v).sum())
    assert nn == po.shape[0] - n_nodes > 0
    pg, qg, wg = [t.cpu().numpy() for t in wf.getNodes()]
    assert same_bits(pg, po) and same_bits(wg, wo)
    assert same_bits(qg[:n_nodes], qo[:n_nodes])
    assert np.allclose(qg[n_nodes:], qo[n_nodes:], atol=2e-6)  # calcDQB tolerance of the blend tests (exp in double)
    # the re-indexed field answers queries like a freshly initialised one with the same nodes
    q = (v[::37] + np.float32(0.003)).astype(np.float32)
    idx_o, _ = oracle.knn(po, q)
    assert np.array_equal(wf.findNeighborsIndex(8, dev(q)).cpu().numpy(), idx_o)
    # a second update with the same frame: the new nodes (dg_w = 2 eps = 2.5 cm, 5 cm grid) leave some gaps
    po2, _, _ = oracle.warpfield_update(po, qo, wo, 0.0125, v)
    wf.update(dev(v))
    assert wf.numNodes() == po2.shape[0]
    assert same_bits(wf.getNodes()[0].cpu().numpy(), po2)


def test_update_without_unsupported_vertices_is_a_noop(fe, oracle):
    depth, intr, pos, dq, w = update_scene()
    wf = make_wf(pos, dq, w)
    v = (pos[:100] + np.float32(0.001)).astype(np.float32)
    assert wf.update(dev(v)) == (0, 0)
    assert wf.numNodes() == pos.shape[0]


def test_update_keeps_voxel_cache_consistent(fe, oracle):
    """integrate (fills the per-voxel 8-NN cache) -> update (adds nodes, selective invalidation) -> integrate:
    the volume must equal the one a fresh warp field with the same nodes produces."""
    import dynfu_b200
    dim = 128
    intr = synth.intr_for(320, 240)
    depth, _, pos, dq, w = update_scene(400)
    rng = np.random.default_rng(2)
    dq = synth.translations_to_dq(rng.normal(0, 0.004, (pos.shape[0], 3)).astype(np.float32))
    v = frame_vertices(oracle, depth, intr)
    dists = dynfu_b200.compute_dists(dev(depth.view(np.int16), torch.int16), intr)
    pose = torch.eye(4, dtype=torch.float64)
    pose[:3, 3] = torch.tensor([-1.5, -1.5, 0.5], dtype=torch.float64)

    def volume():
        vol = dynfu_b200.TsdfVolume((dim, dim, dim))
        vol.setPose(pose)
        vol.setTruncDist(0.06)
        return vol

    cam = torch.eye(4, dtype=torch.float64)
    wf = make_wf(pos, dq, w)
    vol_a = volume()
    vol_a.integrate(dists, cam, intr, wf)            # warm cache with the old nodes
    pool, built0 = wf.cacheStats()
    nu, nn = wf.update(dev(v[::3]))
    assert nn > 0
    _, built1 = wf.cacheStats()
    assert pool > 0 and 0 < built1 < built0          # only the bricks the new nodes can reach were dropped
    vol_a.integrate(dists, cam, intr, wf)            # partially invalidated cache + new nodes
    assert wf.cacheStats()[1] >= built0
    p2, q2, w2 = wf.getNodes()
    wf_b = dynfu_b200.Warpfield()
    wf_b.init(0.0125, dev(pos), dev(dq), dev(w))
    vol_b = volume()
    vol_b.integrate(dists, cam, intr, wf_b)
    wf_c = dynfu_b200.Warpfield()
    wf_c.init(0.0125, p2, q2, w2)
    vol_b.integrate(dists, cam, intr, wf_c)          # everything recomputed from scratch
    assert torch.equal(vol_a.data, vol_b.data)


# ------------------------------------------------------------------------------ the widened frame loop, end to end
def test_process_frame_against_oracle_chain(fe, oracle):
    """DynFusion.processFrame on two depth frames (the second one bulged): depth -> points -> canonical frame /
    nodes -> warp -> correspondences -> solve -> Warpfield::update, replayed step by step with the oracle."""
    import dynfu_b200 as dfu
    from oracle import pyoracle

    rows, cols = 120, 160
    intr = synth.intr_for(cols, rows)
    d0 = synth.sphere_depth(rows, cols, intr)
    d1 = synth.sphere_depth(rows, cols, intr, radius=0.51)
    kp = dfu.KinFuParams(cols=cols, rows=rows, intr=tuple(float(x) for x in intr), volume_dims=(64, 64, 64))
    prm = dfu.DynFuParams(kinfuParams=kp, epsilon=0.03, lambda_=200.0, node_step=16,
                          solver=dfu.CombinedSolverParameters(numIter=3, nonLinearIter=1, linearIter=10, earlyOut=False,
                                                              pcgTolerance=0.0))
    df = dfu.DynFusion(prm)
    assert df.processFrame(torch.from_numpy(d0.view(np.int16)).pin_memory()) is False
    n0 = df.warpfield.numNodes()
    assert df.processFrame(torch.from_numpy(d1.view(np.int16)).pin_memory()) is True
    torch.cuda.synchronize()

    # ---- oracle replay
    xf = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 1.5, 1.5, -0.5], np.float32)
    p0, m0 = oracle.points_normals(d0, intr)
    canon_v, canon_n = oracle.compact_points(p0, m0, xf)
    assert same_bits(df.canonicalVertices.cpu().numpy(), canon_v)
    pos = canon_v[::16].copy()
    assert n0 == pos.shape[0]
    dq = synth.identity_dq(n0)
    w = np.full(n0, np.float32(3) * np.float32(0.03), np.float32)
    p1, m1 = oracle.points_normals(d1, intr)
    live_v, _ = oracle.compact_points(p1, m1, xf)
    assert same_bits(df.liveVertices.cpu().numpy(), live_v)
    warped_v, warped_n = oracle.warp(pos, dq, w, canon_v, canon_n)
    corr_v, _, _, _ = oracle.find_corresponding(warped_v, warped_n, live_v)
    prm_o = pyoracle.default_params(num_iter=3, nonlinear_iter=1, linear_iter=10, lambda_=200.0, pcg_tol=0.0, early_out=0)
    t_o, dq_o, st_o = oracle.solve(pos, dq, w, corr_v, live_v, prm_o)
    st = df.solver.getStats()
    assert abs(st["final_energy"] - st_o[1]) <= 1e-4 * st_o[1]
    assert st_o[1] < st_o[0]  # the solve did reduce the energy
    po, qo, wo = oracle.warpfield_update(pos, dq_o, w, 0.03, warped_v)
    pg, qg, wg = [t.cpu().numpy() for t in df.warpfield.getNodes()]
    assert pg.shape == po.shape
    assert same_bits(pg, po) and same_bits(wg, wo)
    assert np.max(np.abs(qg - qo)) <= 1e-4 * max(np.abs(t_o).max(), 1e-3)


# ------------------------------------------------------------------------------------------------ f4: raycast
def _integrated_sphere_volume(dim=128):
    import dynfu_b200
    intr = synth.intr_for(320, 240)
    depth = synth.sphere_depth(240, 320, intr)
    dists = dynfu_b200.compute_dists(dev(depth.view(np.int16), torch.int16), intr)
    vol = dynfu_b200.TsdfVolume((dim, dim, dim))
    pose = torch.eye(4, dtype=torch.float64)
    pose[:3, 3] = torch.tensor([-1.5, -1.5, 0.5], dtype=torch.float64)
    vol.setPose(pose)
    vol.setTruncDist(0.06)
    for _ in range(3):
        vol.integrate(dists, torch.eye(4, dtype=torch.float64), intr)
    return vol, depth, intr


@pytest.mark.parametrize("yaw", [0.0, 0.2])
def test_raycast_bit_exact(fe, oracle, yaw):
    """TsdfVolume::raycast (tsdf_volume.cu:126-386) of a volume fused from a sphere: points / depth / normals bit-exact"""
    vol, depth, intr = _integrated_sphere_volume()
    cam = np.eye(4)
    cam[:3, :3] = [[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]]
    cam[:3, 3] = [0.4 * np.sin(yaw) * 2, 0.0, 0.05 * yaw]
    pose = vol.getPose().numpy()
    cam2vol = np.linalg.inv(pose) @ cam
    rinv = np.linalg.inv(cam2vol[:3, :3])
    c2v = np.concatenate([cam2vol[:3, :3].reshape(-1), cam2vol[:3, 3]]).astype(np.float32)
    v_host = vol.data.cpu().numpy().view(np.uint32)
    p_o, n_o, d_o = oracle.raycast(v_host, vol.getVoxelSize(), vol.getTruncDist(), c2v, rinv.astype(np.float32).reshape(-1), intr, 240, 320,
                                   want_depth=True)
    p_g, n_g = vol.raycast(torch.as_tensor(cam), intr, 240, 320, want="points")
    d_g, n_g2 = vol.raycast(torch.as_tensor(cam), intr, 240, 320, want="depth")
    p_g, n_g, n_g2 = p_g.cpu().numpy(), n_g.cpu().numpy(), n_g2.cpu().numpy()
    hit = ~np.isnan(p_o[..., 0])
    assert hit.sum() > 5000
    assert np.array_equal(hit, ~np.isnan(p_g[..., 0])) and np.array_equal(hit, ~np.isnan(n_g[..., 0]))
    assert same_bits(p_g[hit], p_o[hit]) and same_bits(n_g[hit], n_o[hit]) and same_bits(n_g2[hit], n_o[hit])
    assert np.array_equal(d_g.cpu().numpy().view(np.uint16), d_o)
    if yaw == 0.0:  # the fused model reproduces the depth image it was fused from to within a voxel (23 mm at 128^3),
        m = hit & (depth > 0)  # except at grazing angles near the silhouette
        err = np.abs(d_o[m].astype(np.int32) - depth[m].astype(np.int32))
        assert np.median(err) <= 12 and np.percentile(err, 90) <= 40
        nrm = n_o[hit][:, :3]
        assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-5)


def test_update_with_fewer_than_eight_nodes(fe, oracle):
    """the reference's findNeighbors returns fewer than 8 nodes when the field is that small: same decisions"""
    pos = np.float32([[1.5, 1.5, 1.5], [1.6, 1.5, 1.5], [1.5, 1.62, 1.5]])
    dq = synth.translations_to_dq(np.float32([[0.01, 0, 0], [0, 0.01, 0], [0, 0, 0.01]]))
    w = np.float32([0.08, 0.08, 0.08])
    rng = np.random.default_rng(3)
    v = (np.float32([1.55, 1.55, 1.5]) + rng.uniform(-0.3, 0.3, (500, 3))).astype(np.float32)
    mask_o = oracle.unsupported(pos, w, v)
    assert 0 < mask_o.sum() < 500
    po, qo, wo = oracle.warpfield_update(pos, dq, w, 0.04, v)
    wf = make_wf(pos, dq, w, 0.04)
    assert np.array_equal(wf.getUnsupportedVertices(dev(v), return_mask=True).cpu().numpy(), mask_o)
    nu, nn = wf.update(dev(v))
    assert nu == int(mask_o.sum()) and nn == po.shape[0] - 3
    pg, qg, wg = [t.cpu().numpy() for t in wf.getNodes()]
    assert same_bits(pg, po) and same_bits(wg, wo) and np.allclose(qg, qo, atol=2e-6)


def test_overlapped_frame_schedule_is_bit_identical(fe):
    """DynFusion.frameDevice(overlap=True) runs the integration of frame i on a second stream, concurrently with the point
    pipeline of frame i+1: volume, node transforms and energies equal the sequential schedule bit for bit over six frames"""
    import dynfu_b200 as dfu

    dim = 128
    pos, _, dg_w, t_true = synth.sphere_nodes(1024, 0.025)
    depth0 = synth.sphere_depth()
    canon = synth.backproject(depth0, synth.INTR)[::2]
    depths = [dev(synth.sphere_depth(bump=0.002 * (1 + i % 3)).view(np.int16), torch.int16) for i in range(3)]
    lives = [dev(canon + 0.002 * (1 + i % 3) * np.array([1.0, 0.5, -0.25], np.float32)) for i in range(3)]
    out = {}
    for overlap in (False, True):
        prm = dfu.DynFuParams(kinfuParams=dfu.KinFuParams(volume_dims=(dim, dim, dim)), epsilon=0.025, lambda_=200.0,
                              solver=dfu.CombinedSolverParameters(numIter=3, nonLinearIter=1, linearIter=8, earlyOut=False,
                                                                  pcgTolerance=0.0))
        df = dfu.DynFusion(prm)
        df.init(dev(canon), None, nodes=(dev(pos), dev(synth.identity_dq(1024)), dev(dg_w)))
        df(torch.from_numpy(depth0.view(np.int16)).pin_memory())
        energies = []
        for i in range(6):
            df.frameDevice(depths[i % 3], lives[i % 3], overlap=overlap)
            energies.append(df.solver.getStats()["final_energy"])
        df.frameSync()
        torch.cuda.synchronize()
        out[overlap] = (df.volume.data.cpu().numpy().copy(), df.warpfield.getNodes()[1].cpu().numpy().copy(), energies)
    assert np.array_equal(out[False][0], out[True][0])
    assert same_bits(out[False][1], out[True][1])
    assert out[False][2] == out[True][2]
    assert (out[True][0] >> 16).max() == 7  # rigid frame 0 + six warped frames


def test_dfu_frame_single_call_equals_the_composed_frame(fe):
    """dfu_frame (one C-ABI call per frame) == compute_dists + warpToLive + initializeProblemInstance + solveAll + integrate
    issued one by one: volume and node transforms bit for bit over three frames"""
    import ctypes as C

    import dynfu_b200 as dfu
    from dynfu_b200 import _lib
    from dynfu_b200._lib import check, dptr, lib

    dim = 128
    pos, _, dg_w, _ = synth.sphere_nodes(1024, 0.025)
    depth0 = synth.sphere_depth()
    canon = synth.backproject(depth0, synth.INTR)[::2]
    depths = [dev(synth.sphere_depth(bump=0.002 * (1 + i)).view(np.int16), torch.int16) for i in range(3)]
    lives = [dev(canon + 0.002 * (1 + i) * np.array([1.0, 0.5, -0.25], np.float32)) for i in range(3)]
    prm = dfu.DynFuParams(kinfuParams=dfu.KinFuParams(volume_dims=(dim, dim, dim)), epsilon=0.025, lambda_=200.0,
                          solver=dfu.CombinedSolverParameters(numIter=3, nonLinearIter=1, linearIter=8, earlyOut=False,
                                                              pcgTolerance=0.0))
    out = {}
    for single in (False, True):
        df = dfu.DynFusion(prm)
        df.init(dev(canon), None, nodes=(dev(pos), dev(synth.identity_dq(1024)), dev(dg_w)))
        df(torch.from_numpy(depth0.view(np.int16)).pin_memory())
        if not single:
            for i in range(3):
                df.frameDevice(depths[i], lives[i], overlap=False)
        else:
            kp = prm.kinfuParams
            fp = _lib.FrameParams()
            fp.volume = df.volume._base_ptr().value
            fp.dims[:] = list(df.volume.dims)
            fp.voxel_size[:] = list(df.volume.getVoxelSize())
            fp.trunc_dist = df.volume.getTruncDist()
            fp.max_weight = df.volume.getMaxWeight()
            v2c = torch.linalg.inv(df.camera_pose) @ df.volume.pose
            fp.vol2cam[:] = [float(x) for x in v2c[:3, :3].reshape(-1)] + [float(x) for x in v2c[:3, 3]]
            fp.intr[:] = [float(x) for x in kp.intr]
            fp.rows, fp.cols = kp.rows, kp.cols
            fp.blend_mode = prm.blend_mode
            fp.z0, fp.z1 = 0, dim
            cache = C.c_void_p()
            check(lib.dfu_pointcache_create(C.byref(cache), 0))
            dists = torch.empty((kp.rows, kp.cols), dtype=torch.int16, device="cuda")
            warped = torch.empty_like(df.canonicalVertices)
            for i in range(3):
                check(lib.dfu_frame(df.warpfield.handle, df.solver._h, cache, C.byref(fp), dptr(depths[i]), kp.cols * 2, dptr(dists),
                                    kp.cols * 2, dptr(df.canonicalVertices), 1, dptr(warped), dptr(lives[i]),
                                    df.canonicalVertices.shape[0], None))
            torch.cuda.synchronize()
            check(lib.dfu_pointcache_destroy(cache))
        torch.cuda.synchronize()
        out[single] = (df.volume.data.cpu().numpy().copy(), df.warpfield.getNodes()[1].cpu().numpy().copy())
    assert np.array_equal(out[False][0], out[True][0]) and same_bits(out[False][1], out[True][1])
