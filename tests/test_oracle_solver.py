"""The 8 OptTest cases of the reference (test/opt_optimisation_test.cpp:212-698) as post-conditions on
the CPU oracle's solver: after solveAll, calcDQB(v).transformVertex(v) ~= target within 1e-3 (:94)."""
import numpy as np
import pytest
from scipy.optimize import least_squares

from oracle import pyoracle
from tests import fixtures_opt as fx


def _solve(o, nodes, dq, src, dst):
    dg_w = np.full(len(nodes), fx.DG_W, np.float32)
    t, dq_new, stats = o.solve(nodes, dq, dg_w, src, dst, pyoracle.default_params(**fx.PARAMS))
    return t, dq_new, stats, dg_w


@pytest.mark.parametrize("case", fx.SINGLE_SOLVE_CASES, ids=[c[0] for c in fx.SINGLE_SOLVE_CASES])
def test_single_solve(oracle, case):
    _, nodes, src, dst = case
    t, dq_new, stats, dg_w = _solve(oracle, nodes, fx.identity_dq(len(nodes)), src, dst)
    warped = oracle.warp(nodes, dq_new, dg_w, src)
    assert np.max(np.abs(warped - dst)) <= fx.MAX_ERROR
    assert stats[1] <= stats[0]


def test_warp_twice(oracle):  # :454-527
    nodes = fx.NODES_GROUP1
    t, dq1, _, dg_w = _solve(oracle, nodes, fx.identity_dq(8), fx.WARP_SRC, fx.WARP_T1)
    assert np.max(np.abs(oracle.warp(nodes, dq1, dg_w, fx.WARP_SRC) - fx.WARP_T1)) <= fx.MAX_ERROR
    warped = oracle.warp(nodes, dq1, dg_w, fx.WARP_SRC)  # warpToLive(canonicalFrame)
    t, dq2, _, _ = _solve(oracle, nodes, dq1, warped, fx.WARP_T2)
    # final check on the ORIGINAL canonical vertices (:518-527)
    assert np.max(np.abs(oracle.warp(nodes, dq2, dg_w, fx.WARP_SRC) - fx.WARP_T2)) <= fx.MAX_ERROR


def test_warp_thrice(oracle):  # :530-630
    nodes = fx.NODES_GROUP1
    _, dq1, _, dg_w = _solve(oracle, nodes, fx.identity_dq(8), fx.WARP_SRC, fx.WARP_T1)
    w1 = oracle.warp(nodes, dq1, dg_w, fx.WARP_SRC)
    _, dq2, _, _ = _solve(oracle, nodes, dq1, w1, fx.WARP_T2)
    assert np.max(np.abs(oracle.warp(nodes, dq2, dg_w, fx.WARP_SRC) - fx.WARP_T2)) <= fx.MAX_ERROR
    w2 = oracle.warp(nodes, dq2, dg_w, w1)  # warpToLive(canonicalFrameWarpedToLive) (:586)
    _, dq3, _, _ = _solve(oracle, nodes, dq2, w2, fx.WARP_T3)
    # final check iterates the ONCE-warped frame (:620-629)
    assert np.max(np.abs(oracle.warp(nodes, dq3, dg_w, w1) - fx.WARP_T3)) <= fx.MAX_ERROR


def test_warp_and_reverse(oracle):  # :632-698
    nodes = fx.NODES_GROUP1
    _, dq1, _, dg_w = _solve(oracle, nodes, fx.identity_dq(8), fx.WARP_SRC, fx.WARP_T1)
    assert np.max(np.abs(oracle.warp(nodes, dq1, dg_w, fx.WARP_SRC) - fx.WARP_T1)) <= fx.MAX_ERROR
    _, dq2, _, _ = _solve(oracle, nodes, dq1, fx.WARP_T1, fx.WARP_SRC)
    # :688-697 compares liveFrame (= original source) warped by the field with itself: net warp ~ identity
    assert np.max(np.abs(oracle.warp(nodes, dq2, dg_w, fx.WARP_SRC) - fx.WARP_SRC)) <= fx.MAX_ERROR


def test_fewer_than_8_nodes_is_a_precondition_violation(oracle):  # reference UB at opt_solver.cpp:63-66
    with pytest.raises(ValueError):
        _solve(oracle, fx.NODES_GROUP1[:5], fx.identity_dq(5), fx.WARP_SRC, fx.WARP_T1)


def test_against_independent_nlls(oracle):
    """Ceres stand-in: scipy's trust-region NLLS on the energy.t residuals (well-posed: lambda>0, P>>N)
    must reach the same energy and translations as the oracle's GN/PCG."""
    rng = np.random.default_rng(7)
    N, P = 12, 400
    nodes = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
    dg_w = np.full(N, 0.8, np.float32)
    canon = rng.uniform(-1, 1, (P, 3)).astype(np.float32)
    live = (canon + 0.01 * np.sin(3 * canon[:, [1, 2, 0]])).astype(np.float32)
    prm = pyoracle.default_params(num_iter=1, nonlinear_iter=4, linear_iter=500, lambda_=5.0, pcg_tol=1e-14)
    t, _, stats = oracle.solve(nodes, fx.identity_dq(N), dg_w, canon, live, prm)

    idx, _ = oracle.knn(nodes, canon)
    nidx, _ = oracle.knn(nodes, nodes)
    w = np.array([[oracle.node_weight(nodes[j], dg_w[j], canon[v]) for j in idx[v]] for v in range(P)], np.float64)
    zero = np.zeros((N, 3))
    th = np.array([oracle.tukey(4.652, 1e-2, (live[v] - canon[v])) for v in range(P)], np.float64)
    wreg = np.sqrt(5.0 / (N * 8))

    def residuals(x):
        tt = x.reshape(N, 3)
        rd = np.sqrt(th)[:, None] * (live.astype(np.float64) - canon - np.einsum("vk,vkc->vc", w, tt[idx]))
        rr = wreg * (tt[nidx] - tt[:, None, :])
        return np.concatenate([rd.ravel(), rr.ravel()])

    sol = least_squares(residuals, zero.ravel(), method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-15)
    E_scipy = float(np.sum(sol.fun ** 2))
    assert abs(stats[1] - E_scipy) <= 1e-8 * max(E_scipy, 1e-30)
    assert np.max(np.abs(sol.x.reshape(N, 3) - t)) <= 1e-6
    # orc_energy agrees with the residual vector at t
    assert abs(oracle.energy(nodes, dg_w, canon, live, prm, t, np.zeros_like(t)) - stats[1]) <= 1e-12 * stats[1] + 1e-18
