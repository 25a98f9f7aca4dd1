"""CPU checks of the oracle's front-end rows (f3 points/normals, f2 correspondences).  The reference
holds no tests for these; the oracle is pinned against an independent numpy float32 restatement of the kernel and,
for the correspondences, against the reference's own nanoflann (oracle/_ref)."""
import numpy as np

from tests import synth


def np_points_normals(depth, intr):
    """src/kfusion/cuda/imgproc.cu:187-215 in numpy float32, one rounding per operation"""
    f = np.float32
    rows, cols = depth.shape
    finvx, finvy, cx, cy = f(1) / f(intr[0]), f(1) / f(intr[1]), f(intr[2]), f(intr[3])
    z = depth.astype(np.float32) * f(0.001)
    u = np.arange(cols, dtype=np.float32)[None, :].repeat(rows, 0)
    v = np.arange(rows, dtype=np.float32)[:, None].repeat(cols, 1)

    def reproj(uu, vv, zz):
        return np.stack([zz * (uu - cx) * finvx, zz * (vv - cy) * finvy, zz], -1)

    pts = np.full((rows, cols, 4), np.nan, np.float32)
    nrm = np.full((rows, cols, 4), np.nan, np.float32)
    if rows < 2 or cols < 2:
        return pts, nrm
    z00, z01, z10 = z[:-1, :-1], z[:-1, 1:], z[1:, :-1]
    ok = (z00 * z01 * z10) != 0
    v00 = reproj(u[:-1, :-1], v[:-1, :-1], z00)
    v01 = reproj(u[:-1, 1:], v[:-1, 1:], z01)
    v10 = reproj(u[1:, :-1], v[1:, :-1], z10)
    a, b = v01 - v00, v10 - v00
    c = np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1], a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                  a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], -1)
    with np.errstate(invalid="ignore", divide="ignore"):
        ln = np.sqrt((c[..., 0] * c[..., 0] + c[..., 1] * c[..., 1]) + c[..., 2] * c[..., 2])
        n = -(c / ln[..., None])
    P = np.concatenate([v00, np.zeros_like(z00)[..., None]], -1)
    N = np.concatenate([n, np.zeros_like(z00)[..., None]], -1)
    pts[:-1, :-1][ok] = P[ok]
    nrm[:-1, :-1][ok] = N[ok]
    return pts, nrm


def test_points_normals_against_numpy(oracle):
    rng = np.random.default_rng(0)
    for rows, cols in [(120, 160), (5, 3), (1, 4)]:
        intr = synth.intr_for(cols, rows)
        depth = synth.sphere_depth(rows, cols, intr, bump=0.03).copy()
        depth[rng.random((rows, cols)) < 0.05] = 0
        p_o, n_o = oracle.points_normals(depth, intr)
        p_n, n_n = np_points_normals(depth, intr)
        assert np.array_equal(np.isnan(p_o), np.isnan(p_n))
        assert np.array_equal(np.nan_to_num(p_o, nan=3.0), np.nan_to_num(p_n, nan=3.0))
        assert np.array_equal(np.nan_to_num(n_o, nan=3.0), np.nan_to_num(n_n, nan=3.0))
        assert np.isnan(p_o[-1]).all() and np.isnan(p_o[:, -1]).all()  # last row / column never valid


def test_points_reproject_onto_their_pixel(oracle):
    intr = synth.intr_for(160, 120)
    depth = synth.sphere_depth(120, 160, intr)
    p, n = oracle.points_normals(depth, intr)
    ys, xs = np.nonzero(~np.isnan(p[..., 0]))
    q = p[ys, xs]
    assert np.allclose(q[:, 0] / q[:, 2] * intr[0] + intr[2], xs, atol=1e-3)
    assert np.allclose(q[:, 1] / q[:, 2] * intr[1] + intr[3], ys, atol=1e-3)
    assert np.array_equal(q[:, 2], depth[ys, xs].astype(np.float32) * np.float32(0.001))
    assert (n[ys, xs][:, 2] < 0).all()  # normals of a front-facing surface point towards the camera


def test_compact_points_order_and_transform(oracle):
    intr = synth.intr_for(64, 48)
    depth = synth.sphere_depth(48, 64, intr).copy()
    depth[10:20, 5:9] = 0
    p, n = oracle.points_normals(depth, intr)
    v, m = oracle.compact_points(p, n)
    mask = ~np.isnan(p[..., 0])
    assert np.array_equal(v, p[mask][:, :3]) and np.array_equal(m, n[mask][:, :3])
    xf = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 1.5, 1.5, -0.5], np.float32)
    v2, m2 = oracle.compact_points(p, n, xf)
    assert np.allclose(v2, v + [1.5, 1.5, -0.5], atol=1e-6) and np.array_equal(m2, m)


def test_find_corresponding_brute_equals_reference_kdtree(oracle, oracle_nf):
    rng = np.random.default_rng(4)
    canon = rng.uniform(0, 3, (5000, 3)).astype(np.float32)
    normals = rng.normal(size=(5000, 3)).astype(np.float32)
    live = (canon[rng.integers(0, 5000, 4000)] + rng.normal(0, 0.02, (4000, 3))).astype(np.float32)
    v_b, n_b, i_b, ties = oracle.find_corresponding(canon, normals, live)
    v_k, n_k, i_k, _ = oracle_nf.find_corresponding(canon, normals, live)
    assert ties == 0
    assert np.array_equal(i_b, i_k) and np.array_equal(v_b, v_k) and np.array_equal(n_b, n_k)
    assert np.array_equal(v_b, canon[i_b])


# ---- Warpfield::update (src/dynfu/warp_field.cpp:34-95) -----------------------------------------------------------
def test_unsupported_threshold(oracle):
    pos = np.float32([[0, 0, 0]] + [[10 + i, 0, 0] for i in range(7)])
    w = np.full(8, 0.5, np.float32)
    verts = np.float32([[0.25, 0, 0], [0.5, 0, 0], [0.49999997, 0, 0], [0, 0.6, 0]])
    # ratio 0.5 -> supported; exactly 1 -> unsupported (`min >= 1`, :56); just below 1 -> supported; 1.2 -> unsupported
    assert oracle.unsupported(pos, w, verts).tolist() == [False, True, False, True]


def np_voxel_grid(pts, leaf):
    """independent numpy restatement of pcl::VoxelGrid (PCL 1.8.1 voxel_grid.hpp), ascending index inside a cell"""
    f = np.float32
    inv = f(1) / f(leaf)
    mn, mx = pts.min(0), pts.max(0)
    min_b = np.floor(mn * inv).astype(np.int64)
    div_b = np.floor(mx * inv).astype(np.int64) - min_b + 1
    ijk = (np.floor(pts * inv) - min_b.astype(np.float32)).astype(np.int64)
    key = ijk[:, 0] + ijk[:, 1] * div_b[0] + ijk[:, 2] * div_b[0] * div_b[1]
    out = []
    for k in np.unique(key):
        sel = pts[key == k]
        s = np.zeros(3, np.float32)
        for p in sel:
            s = s + p
        out.append(s / f(len(sel)))
    return np.array(out, np.float32)


def test_voxel_grid_against_numpy(oracle):
    rng = np.random.default_rng(8)
    for lo, hi, n in [(0.2, 1.4, 3000), (-0.7, 0.9, 2500)]:
        pts = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
        c_o = oracle.voxel_grid(pts, 0.05)
        c_n = np_voxel_grid(pts, 0.05)
        assert c_o.shape == c_n.shape and np.array_equal(c_o, c_n)
        # PCL's unstable std::sort only reorders float additions inside a cell
        c_s = oracle.voxel_grid(pts, 0.05, order_mode=1)
        assert c_s.shape == c_o.shape and np.abs(c_s - c_o).max() < 1e-6


def test_voxel_grid_refuses_int32_overflow(oracle):
    assert oracle.voxel_grid(np.float32([[0, 0, 0], [1e4, 1e4, 1e4]]), 0.001) is None  # PCL returns its input


def test_update_appends_nodes(oracle):
    pos, dq, w, _ = synth.sphere_nodes(64, 0.0125, rotations=True)
    rng = np.random.default_rng(1)
    verts = np.concatenate([pos[:20] + np.float32(0.002), pos[:5] + np.float32([0.3, 0.0, 0.0])]).astype(np.float32)
    po, qo, wo = oracle.warpfield_update(pos, dq, w, 0.0125, verts)
    n_new = po.shape[0] - 64
    assert 1 <= n_new <= 5
    assert np.array_equal(po[:64], pos) and np.array_equal(qo[:64], dq) and np.array_equal(wo[:64], w)
    assert np.all(wo[64:] == np.float32(2 * 0.0125))                       # dg_w = 2 * epsilon (:80)
    assert np.array_equal(qo[64:], oracle.blend(pos, dq, w, po[64:]))      # calcDQB against the OLD nodes (:79)
    # supported-only frames add nothing
    po2, _, _ = oracle.warpfield_update(pos, dq, w, 0.0125, verts[:20])
    assert po2.shape[0] == 64


# ---- TsdfVolume::raycast (src/kfusion/cuda/tsdf_volume.cu:126-386) -----------------------------------------------------
def test_raycast_finds_an_analytic_sphere(oracle):
    dim, size, trunc = 64, 3.0, 0.15
    vs = np.float32(size / dim)
    idx = np.arange(dim, dtype=np.float32) * vs
    Z, Y, X = np.meshgrid(idx, idx, idx, indexing="ij")
    c = np.float32([1.5, 1.5, 1.5])
    sdf = np.sqrt((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2) - np.float32(0.5)
    tsdf = np.clip(sdf / trunc, -1, 1).astype(np.float16)
    vol = tsdf.view(np.uint16).astype(np.uint32) | (np.uint32(1) << 16)
    intr = synth.intr_for(160, 120)
    cam2vol = np.float32([1, 0, 0, 0, 1, 0, 0, 0, 1, 1.5, 1.5, -0.5])  # camera 2 m in front of the sphere centre
    rinv = np.eye(3, dtype=np.float32).reshape(-1)
    pts, nrm, dep = oracle.raycast(vol, [vs, vs, vs], trunc, cam2vol, rinv, intr, 120, 160, want_depth=True)
    hit = ~np.isnan(pts[..., 0])
    assert 1500 < hit.sum() < 120 * 160
    p = pts[hit][:, :3] + np.float32([1.5, 1.5, -0.5]) - c  # camera frame -> relative to the sphere centre
    r = np.linalg.norm(p, axis=1)
    # zero crossing located to a fraction of a voxel -- except on grazing rays at the silhouette, where the reference's
    # nearest-voxel sign test fires early and the linear refinement extrapolates (its behaviour, reproduced as is)
    good = np.abs(r - 0.5) < 0.5 * vs
    assert good.mean() > 0.95
    n = nrm[hit][:, :3]
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-5)
    assert (np.sum(n * (p / r[:, None]), axis=1)[good] > 0.98).all()  # gradient of the TSDF = outward radial direction
    assert np.array_equal(dep[hit], np.clip((pts[hit][:, 2] * np.float32(1000)).astype(np.int32), 0, 65535).astype(np.uint16))
    assert (dep[~hit] == 0).all() and np.isnan(nrm[~hit]).all()
    # rays that miss the volume entirely
    far = np.float32([1, 0, 0, 0, 1, 0, 0, 0, 1, 50.0, 1.5, -0.5])
    pts2, _ = oracle.raycast(vol, [vs, vs, vs], trunc, far, rinv, intr, 12, 16)
    assert np.isnan(pts2).all()
