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
