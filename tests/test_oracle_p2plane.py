"""The north-star's point-to-plane SE(3) data term (BASELINE.json north_star (4)) has no reference implementation: the
double-precision oracle is the yardstick, and this file pins the oracle itself against scipy.optimize.least_squares on
the same nonlinear energy."""
import numpy as np
import pytest

from oracle import pyoracle
from tests import synth


def rigid_scene(n_nodes=16, n_pts=400, seed=5, angle=0.06, shift=(0.01, -0.008, 0.012)):
    """points on a bumpy patch, moved by ONE rigid transform: every node increment should recover it"""
    rng = np.random.default_rng(seed)
    gx, gy = np.meshgrid(np.linspace(1.2, 1.8, 4), np.linspace(1.2, 1.8, n_nodes // 4))
    pos = np.stack([gx.ravel(), gy.ravel(), 1.5 + 0.05 * np.sin(5 * gx.ravel())], 1).astype(np.float32)
    pos += rng.normal(0, 0.003, pos.shape).astype(np.float32)
    xy = rng.uniform(1.2, 1.8, (n_pts, 2))
    z = 1.5 + 0.05 * np.sin(5 * xy[:, 0]) + 0.04 * np.cos(4 * xy[:, 1])
    canon = np.concatenate([xy, z[:, None]], 1).astype(np.float32)
    # surface normals of z = f(x, y)
    nrm = np.stack([-0.25 * np.cos(5 * xy[:, 0]), 0.16 * np.sin(4 * xy[:, 1]), np.ones(n_pts)], 1)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    ax = np.array([0.3, -0.5, 0.8]); ax /= np.linalg.norm(ax)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K
    c0 = np.array([1.5, 1.5, 1.5])
    live = (canon - c0) @ R.T + c0 + np.array(shift)
    live_n = nrm @ R.T
    dg_w = np.full(len(pos), 0.25, np.float32)
    return pos, dg_w, canon, live.astype(np.float32), live_n.astype(np.float32), R, c0 + np.array(shift) - R @ c0


def test_p2plane_recovers_a_rigid_motion(oracle):
    pos, dg_w, canon, live, live_n, R, t = rigid_scene()
    prm = pyoracle.default_params(num_iter=6, nonlinear_iter=3, linear_iter=200, lambda_=50.0, psi_data=1.0, pcg_tol=1e-12)
    X, dq, st = oracle.solve_p2plane(pos, synth.identity_dq(len(pos)), dg_w, canon, live, live_n, prm)
    assert st[1] < 1e-6 * st[0]
    # every increment maps the surface onto the live surface: compare the warped points, not the (gauge-free) transforms
    Xr = X[:, :9].reshape(-1, 3, 3)
    moved = np.einsum("nij,pj->npi", Xr, canon.astype(np.float64)) + X[:, None, 9:]
    target = canon.astype(np.float64) @ R.T + t
    d = np.einsum("npi,pi->np", moved - target[None], live_n.astype(np.float64))
    idx, _ = oracle.knn(pos, canon)
    near = np.take_along_axis(np.abs(d).T, idx[:, :1].astype(np.int64), axis=1)  # residual under each point's nearest node
    assert near.max() < 5e-4


def test_p2plane_against_scipy_least_squares(oracle):
    """same nonlinear energy (fixed Tukey weights at the start), minimised by an independent LM implementation"""
    scipy_opt = pytest.importorskip("scipy.optimize")
    from scipy.spatial.transform import Rotation

    pos, dg_w, canon, live, live_n, _, _ = rigid_scene(n_nodes=12, n_pts=150, angle=0.03, shift=(0.004, 0.0, -0.003))
    N = len(pos)
    prm = pyoracle.default_params(num_iter=1, nonlinear_iter=8, linear_iter=300, lambda_=20.0, psi_data=1.0, pcg_tol=1e-13)
    X, _, st = oracle.solve_p2plane(pos, synth.identity_dq(N), dg_w, canon, live, live_n, prm)
    ident = np.tile(np.concatenate([np.eye(3).ravel(), np.zeros(3)]), (N, 1))

    def to_X(x):
        x = x.reshape(N, 6)
        Rm = Rotation.from_rotvec(x[:, :3]).as_matrix()
        return np.concatenate([Rm.reshape(N, 9), x[:, 3:]], 1)

    # residual vector whose squared norm is the oracle's energy (theta fixed at the identity, like num_iter = 1)
    idx, _ = oracle.knn(pos, canon)
    w = np.array([[oracle.node_weight(pos[j], dg_w[j], canon[v]) for j in idx[v]] for v in range(len(canon))], np.float64)
    wn = w / w.sum(1, keepdims=True)
    nn_idx, _ = oracle.knn(pos, pos)
    theta = np.array([oracle.tukey(4.652, 1.0, live[v] - canon[v]) for v in range(len(canon))], np.float64)
    wreg = np.sqrt(20.0 / (N * 8))

    def residuals(x):
        Xm = to_X(x)
        Rm, tm = Xm[:, :9].reshape(N, 3, 3), Xm[:, 9:]
        q = np.einsum("pkij,pj->pki", Rm[idx], canon.astype(np.float64)) + tm[idx]
        p = (wn[..., None] * q).sum(1)
        r_data = np.sqrt(theta) * np.einsum("pi,pi->p", live_n.astype(np.float64), p - live)
        r_reg = []
        for n in range(N):
            for m in nn_idx[n]:
                if m != n:
                    g = pos[m].astype(np.float64)
                    r_reg.append(wreg * ((Rm[n] @ g + tm[n]) - (Rm[m] @ g + tm[m])))
        return np.concatenate([r_data, np.concatenate(r_reg)])

    x0 = np.zeros(6 * N)
    assert abs(np.sum(residuals(x0) ** 2) - oracle.energy_p2plane(pos, dg_w, canon, live, live_n, prm, ident)) < 1e-12 + 1e-9 * st[0]
    sol = scipy_opt.least_squares(residuals, x0, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15)
    E_scipy = float(np.sum(sol.fun ** 2))
    E_oracle = oracle.energy_p2plane(pos, dg_w, canon, live, live_n, prm, X, ident)
    assert abs(E_oracle - st[1]) <= 1e-9 * max(st[1], 1e-30)
    assert E_oracle <= E_scipy * (1 + 1e-4) + 1e-15 and E_scipy <= E_oracle * (1 + 1e-4) + 1e-15, (E_oracle, E_scipy, st)


def test_p2plane_robust_regularisation_against_scipy(oracle):
    """reg_mode 1 (DynamicFusion eq. 8 as IRLS: edge weight w_reg^2 max(dg_w_i, dg_w_j) h_ij, h = Huber weight): with the
    Tukey and Huber weights frozen at the result X1 of the first outer iteration, the second outer iteration of the oracle
    must reach the minimum an independent LM finds for the same weighted residual vector"""
    scipy_opt = pytest.importorskip("scipy.optimize")
    from scipy.spatial.transform import Rotation

    pos, dg_w, canon, live, live_n, _, _ = rigid_scene(n_nodes=12, n_pts=150, angle=0.03, shift=(0.004, 0.0, -0.003))
    rng = np.random.default_rng(3)
    dg_w = (dg_w + rng.uniform(0.0, 0.08, dg_w.shape)).astype(np.float32)          # alpha_ij varies
    live = (live + 0.004 * np.sin(9.0 * canon[:, :1]) * live_n).astype(np.float32)  # not a rigid motion: edges disagree
    N = len(pos)
    ident = np.tile(np.concatenate([np.eye(3).ravel(), np.zeros(3)]), (N, 1))
    lam = 20.0
    kw = dict(nonlinear_iter=10, linear_iter=300, lambda_=lam, psi_data=1.0, pcg_tol=1e-13, reg_mode=1)
    X1, _, _ = oracle.solve_p2plane(pos, synth.identity_dq(N), dg_w, canon, live, live_n, pyoracle.default_params(num_iter=1, **kw))
    nn_idx, _ = oracle.knn(pos, pos)
    R1, t1 = X1[:, :9].reshape(N, 3, 3), X1[:, 9:]
    edges = [(n, m) for n in range(N) for m in nn_idx[n] if m != n]
    r1 = np.array([np.linalg.norm((R1[n] @ pos[m].astype(np.float64) + t1[n]) - (R1[m] @ pos[m].astype(np.float64) + t1[m]))
                   for n, m in edges])
    psi = float(np.median(r1))
    assert r1.min() < psi < r1.max()  # both Huber branches occur
    prm = pyoracle.default_params(num_iter=2, psi_reg=psi, **kw)
    X2, _, st = oracle.solve_p2plane(pos, synth.identity_dq(N), dg_w, canon, live, live_n, prm)

    idx, _ = oracle.knn(pos, canon)
    w = np.array([[oracle.node_weight(pos[j], dg_w[j], canon[v]) for j in idx[v]] for v in range(len(canon))], np.float64)
    wn = w / w.sum(1, keepdims=True)
    q1 = np.einsum("pkij,pj->pki", R1[idx], canon.astype(np.float64)) + t1[idx]
    p1 = (wn[..., None] * q1).sum(1)
    theta = np.array([oracle.tukey(4.652, 1.0, (live[v] - p1[v]).astype(np.float32)) for v in range(len(canon))], np.float64)
    h = np.where(r1 <= np.float32(psi), 1.0, np.float64(np.float32(psi)) / r1)
    alpha = np.array([max(dg_w[n], dg_w[m]) for n, m in edges], np.float64)
    we = np.sqrt(lam / (N * 8) * alpha * h)

    def to_X(x):
        x = x.reshape(N, 6)
        return np.concatenate([Rotation.from_rotvec(x[:, :3]).as_matrix().reshape(N, 9), x[:, 3:]], 1)

    def residuals(x):
        Xm = to_X(x)
        Rm, tm = Xm[:, :9].reshape(N, 3, 3), Xm[:, 9:]
        q = np.einsum("pkij,pj->pki", Rm[idx], canon.astype(np.float64)) + tm[idx]
        p = (wn[..., None] * q).sum(1)
        r_data = np.sqrt(theta) * np.einsum("pi,pi->p", live_n.astype(np.float64), p - live)
        r_reg = [we[i] * ((Rm[n] @ pos[m].astype(np.float64) + tm[n]) - (Rm[m] @ pos[m].astype(np.float64) + tm[m]))
                 for i, (n, m) in enumerate(edges)]
        return np.concatenate([r_data, np.concatenate(r_reg)])

    x0 = np.zeros(6 * N)
    E_id = oracle.energy_p2plane(pos, dg_w, canon, live, live_n, prm, ident, X1)
    assert abs(np.sum(residuals(x0) ** 2) - E_id) < 1e-12 + 1e-7 * E_id  # same weights, same energy
    sol = scipy_opt.least_squares(residuals, x0, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15)
    E_scipy = float(np.sum(sol.fun ** 2))
    E_oracle = oracle.energy_p2plane(pos, dg_w, canon, live, live_n, prm, X2, X1)
    assert abs(E_oracle - st[1]) <= 1e-9 * max(st[1], 1e-30)
    assert E_oracle <= E_scipy * (1 + 1e-4) + 1e-15 and E_scipy <= E_oracle * (1 + 1e-4) + 1e-15, (E_oracle, E_scipy, st)
    # and the robust weights matter: the plain quadratic regulariser ends somewhere else
    X2q, _, _ = oracle.solve_p2plane(pos, synth.identity_dq(N), dg_w, canon, live, live_n,
                                     pyoracle.default_params(num_iter=2, psi_reg=psi, **{**kw, "reg_mode": 0}))
    assert np.max(np.abs(X2q - X2)) > 1e-5


def test_p2plane_robust_regularisation_limits(oracle):
    """reg_mode 1 with a threshold no edge exceeds and one common dg_w is the quadratic regulariser with lambda scaled by
    alpha = dg_w (every Huber weight is 1); and the Huber weight itself is calcHuberWeight (opt_solver.cpp:233-239)"""
    pos, dg_w, canon, live, live_n, _, _ = rigid_scene(n_nodes=16, n_pts=300, angle=0.04)
    live = (live + 0.003 * np.sin(9.0 * canon[:, :1]) * live_n).astype(np.float32)
    N = len(pos)
    assert np.all(dg_w == dg_w[0])
    kw = dict(num_iter=2, nonlinear_iter=4, linear_iter=200, psi_data=1.0, pcg_tol=1e-13)
    Xr, _, str_ = oracle.solve_p2plane(pos, synth.identity_dq(N), dg_w, canon, live, live_n,
                                       pyoracle.default_params(lambda_=40.0, psi_reg=1e3, reg_mode=1, **kw))
    Xq, _, stq = oracle.solve_p2plane(pos, synth.identity_dq(N), dg_w, canon, live, live_n,
                                      pyoracle.default_params(lambda_=40.0 * float(dg_w[0]), psi_reg=1e3, reg_mode=0, **kw))
    assert np.max(np.abs(Xr - Xq)) <= 1e-9 and abs(str_[1] - stq[1]) <= 1e-9 * stq[1]
    assert oracle.lib.orc_huber(np.float32(0.5), np.float32(0.25)) == 1.0
    assert abs(oracle.lib.orc_huber(np.float32(0.5), np.float32(-2.0)) - 0.25) < 1e-7
