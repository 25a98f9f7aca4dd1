"""Synthetic inputs for the parity tests and bench.py.  Pure numpy; seeds fixed.

Camera = the reference's defaults (src/kfusion/kinfu.cpp:16-25): 640x480, fx=fy=525, cx=319.5, cy=239.5,
3 m cubic volume, volume_pose = translate(-1.5,-1.5,0.5), camera pose identity, trunc 0.04 m, max_weight 64.
Deformation nodes live in VOLUME-LOCAL metres (the frame marching cubes emits)."""
import numpy as np

SEED = 1234
INTR = np.array([525.0, 525.0, 319.5, 239.5], np.float32)
VOLUME_SIZE = 3.0
VOLUME_T = np.array([-1.5, -1.5, 0.5], np.float32)  # volume_pose translation (camera pose = identity)
# vol2cam = inv(camera_pose) * volume_pose  (src/kfusion/tsdf_volume.cpp:83): row-major R then t
VOL2CAM = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, -1.5, -1.5, 0.5], np.float32)
TRUNC = 0.04
MAX_WEIGHT = 64


def intr_for(cols, rows):
    if (cols, rows) == (640, 480):
        return INTR.copy()
    if (cols, rows) == (1280, 720):
        return np.array([1050.0, 1050.0, 639.5, 359.5], np.float32)
    s = cols / 640.0
    return np.array([525.0 * s, 525.0 * s, cols / 2 - 0.5, rows / 2 - 0.5], np.float32)


def voxel_size(dim):
    return np.full(3, np.float32(VOLUME_SIZE) / np.float32(dim), np.float32)


def _rays(rows, cols, intr):
    u, v = np.meshgrid(np.arange(cols, dtype=np.float64), np.arange(rows, dtype=np.float64))
    return np.stack([(u - intr[2]) / intr[0], (v - intr[3]) / intr[1], np.ones_like(u)], -1)


def sphere_depth(rows=480, cols=640, intr=None, center=(0.0, 0.0, 2.0), radius=0.5, bump=0.0):
    """Depth image (uint16 mm, z-depth like a Kinect) of a sphere by analytic ray intersection.
    bump>0 renders the radially displaced 'live' sphere r(dir) = radius + bump*sin(3*theta)."""
    intr = intr_for(cols, rows) if intr is None else intr
    d = _rays(rows, cols, intr)
    c = np.asarray(center, np.float64)
    a = np.sum(d * d, -1)
    b = -2.0 * d @ c
    r = radius
    for _ in range(4 if bump else 1):
        disc = b * b - 4 * a * (c @ c - r * r)
        hit = disc >= 0
        s = np.where(hit, (-b - np.sqrt(np.where(hit, disc, 0))) / (2 * a), 0.0)
        if bump:
            pt = d * s[..., None] - c
            theta = np.arctan2(pt[..., 1], pt[..., 0])
            r = radius + bump * np.sin(3 * theta)
    z = s * d[..., 2]
    return np.where(hit, np.clip(np.round(z * 1000.0), 0, 65535), 0).astype(np.uint16)


def cylinder_depth(rows=480, cols=640, intr=None, center=(0.0, 0.0, 2.0), radius=0.3, length=1.6, kappa=0.0):
    """Depth (uint16 mm) of a cylinder whose axis is parallel to y, optionally bent by x += kappa*(y-cy)^2
    (the axis becomes a parabola; solved per ray by fixed-point iteration from the straight cylinder)."""
    intr = intr_for(cols, rows) if intr is None else intr
    d = _rays(rows, cols, intr)
    c = np.asarray(center, np.float64)
    a = d[..., 0] ** 2 + d[..., 2] ** 2
    shift = np.zeros(d.shape[:2])
    for _ in range(8 if kappa else 1):
        cx = c[0] + shift
        b = -2.0 * (d[..., 0] * cx + d[..., 2] * c[2])
        cc = cx ** 2 + c[2] ** 2 - radius * radius
        disc = b * b - 4 * a * cc
        hit = disc >= 0
        s = np.where(hit, (-b - np.sqrt(np.where(hit, disc, 0))) / (2 * a), 0.0)
        y = s * d[..., 1]
        shift = np.where(hit, kappa * (y - c[1]) ** 2, shift)
    hit &= np.abs(y - c[1]) <= length / 2
    z = s * d[..., 2]
    return np.where(hit, np.clip(np.round(z * 1000.0), 0, 65535), 0).astype(np.uint16)


def _jitter(rng, shape):
    return rng.uniform(-1e-4, 1e-4, shape)


def sphere_nodes(n, epsilon, center_cam=(0.0, 0.0, 2.0), radius=0.5, seed=SEED, rotations=False):
    """n nodes on a Fibonacci lattice on the sphere, VOLUME-LOCAL coordinates, dg_w = 3*epsilon
    (dyn_fusion.cpp:156-158), DQ = identity rotation + t_i = 0.02*sin(3*theta_i)*n_i + 0.01*x
    (optionally small random rotations <= 5 deg)."""
    rng = np.random.default_rng(seed)
    i = np.arange(n, dtype=np.float64) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    th = np.pi * (1 + 5 ** 0.5) * i
    nrm = np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], -1)
    c_vol = np.asarray(center_cam, np.float64) - VOLUME_T
    pos = c_vol + radius * nrm + _jitter(rng, (n, 3))
    theta = np.arctan2(nrm[:, 1], nrm[:, 0])
    t = 0.02 * np.sin(3 * theta)[:, None] * nrm + np.array([0.01, 0, 0])
    dq = translations_to_dq(t, rng if rotations else None)
    return pos.astype(np.float32), dq, np.full(n, 3 * epsilon, np.float32), t.astype(np.float32)


def cylinder_nodes(n_theta, n_y, epsilon, center_cam=(0.0, 0.0, 2.0), radius=0.3, length=1.6, seed=SEED,
                   front_only=True):
    """(n_theta x n_y) theta-y grid of nodes on the cylinder (camera-facing half when front_only)."""
    rng = np.random.default_rng(seed)
    span = np.pi if front_only else 2 * np.pi
    th = (np.arange(n_theta) + 0.5) / n_theta * span + (np.pi if front_only else 0.0)  # z = c + r sin(th) < c
    yy = (np.arange(n_y) + 0.5) / n_y * length - length / 2
    T, Y = np.meshgrid(th, yy, indexing="ij")
    c_vol = np.asarray(center_cam, np.float64) - VOLUME_T
    pos = np.stack([c_vol[0] + radius * np.cos(T), c_vol[1] + Y, c_vol[2] + radius * np.sin(T)], -1).reshape(-1, 3)
    pos = pos + _jitter(rng, pos.shape)
    n = pos.shape[0]
    return pos.astype(np.float32), identity_dq(n), np.full(n, 3 * epsilon, np.float32)


def identity_dq(n):
    dq = np.zeros((n, 8), np.float32)
    dq[:, 0] = 1
    return dq


def translations_to_dq(t, rng=None, max_deg=5.0):
    """DQ(rot, t) as the reference builds it (dual_quaternion.hpp:42-45): dual = 0.5*(0,t)*real."""
    t = np.asarray(t, np.float64)
    n = t.shape[0]
    if rng is None:
        q = np.zeros((n, 4))
        q[:, 0] = 1
    else:
        axis = rng.normal(size=(n, 3))
        axis /= np.linalg.norm(axis, axis=1, keepdims=True)
        ang = np.deg2rad(rng.uniform(0, max_deg, n))
        q = np.concatenate([np.cos(ang / 2)[:, None], np.sin(ang / 2)[:, None] * axis], 1)
    a = np.concatenate([np.zeros((n, 1)), t], 1)
    dual = 0.5 * _qmul(a, q)
    return np.concatenate([q, dual], 1).astype(np.float32)


def _qmul(p, q):
    a, b, c, d = p.T
    ar, br, cr, dr = q.T
    return np.stack([a * ar - b * br - c * cr - d * dr, a * br + b * ar + c * dr - d * cr,
                     a * cr - b * dr + c * ar + d * br, a * dr + b * cr - c * br + d * ar], -1)


def backproject(depth, intr, stride=1):
    """Valid depth pixels -> points in VOLUME-LOCAL metres (camera pose identity)."""
    rows, cols = depth.shape
    v, u = np.mgrid[0:rows:stride, 0:cols:stride]
    z = depth[::stride, ::stride].astype(np.float64) * 1e-3
    m = z > 0
    x = (u[m] - intr[2]) / intr[0] * z[m]
    y = (v[m] - intr[3]) / intr[1] * z[m]
    p_cam = np.stack([x, y, z[m]], -1)
    return (p_cam - VOLUME_T).astype(np.float32)


def bend(points_vol, kappa, center_cam=(0.0, 0.0, 2.0)):
    """C3's bend: x += kappa * y^2 (y relative to the cylinder centre)."""
    c_vol = np.asarray(center_cam, np.float64) - VOLUME_T
    p = points_vol.astype(np.float64).copy()
    p[:, 0] += kappa * (p[:, 1] - c_vol[1]) ** 2
    return p.astype(np.float32)


def assert_no_knn_ties(oracle, nodes, queries):
    _, ties = oracle.knn(nodes, queries)
    assert ties == 0, f"{ties} queries have bit-equal distances among their 9 nearest nodes"


def with_wall(depth, wall_mm=3000):
    """the same depth image in front of a wall perpendicular to the optical axis at `wall_mm`: every pixel has a depth
    (the dense-depth variant of the integration workload: the whole frustum in front of the surfaces is updated)"""
    d = depth.copy()
    d[d == 0] = wall_mm
    return d
