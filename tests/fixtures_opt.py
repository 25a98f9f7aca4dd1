"""Fixtures of the reference's solver tests (test/opt_optimisation_test.cpp), verbatim values.
Shared by the CPU-oracle tests and the GPU parity tests so both read like the reference's OptTest."""
import numpy as np

# :56-63
NODES_GROUP1 = np.array([(3, 1, -1), (1, 1, 1), (-1, 2, 3), (-1, -1, 1), (-2, -1, -1), (2, -1, -3), (-1, 1, -1),
                         (2, 1, 1)], np.float32)
# :66-75
NODES_GROUP2 = np.array([(10, 10, 10), (9, 11.1, 10), (10, 9, 10), (10, 12, 9), (9, 11, 10), (12, 10, 9), (9, 9, 12),
                         (10.5, 9, 9), (10.5, 12, 12), (11, 11, 10.9)], np.float32)
ALL_NODES = np.concatenate([NODES_GROUP1, NODES_GROUP2])  # :78-79
DG_W = 2.0  # :53
MAX_ERROR = 1e-3  # :94
EPSILON_DYNFU = 0.0015  # :113
# :115-122, :38-44
PARAMS = dict(num_iter=32, nonlinear_iter=16, linear_iter=256, tukey_offset=4.652, psi_data=1e-2, lambda_=0.0,
              psi_reg=1e-4)


def identity_dq(n):
    dq = np.zeros((n, 8), np.float32)
    dq[:, 0] = 1.0
    return dq


def _diag(vals):
    return np.array([(v, v, v) for v in vals], np.float32)


# (name, nodes, source vertices, target vertices) of the single-solve cases
SINGLE_SOLVE_CASES = [
    ("SingleVertexOneGroup", NODES_GROUP1, np.array([(0, 0.04, 0)], np.float32),
     np.array([(0.01, 0.03, 0)], np.float32)),  # :212
    ("TwoVerticesOneNotMoving", ALL_NODES, np.array([(0, 0.05, 1), (2, 2, 2)], np.float32),
     np.array([(0.01, 0.04, 1.01), (2, 2, 2)], np.float32)),  # :243
    ("MultipleVerticesOneGroup", NODES_GROUP1, _diag([-3, -2, 0.01, 2, 3]), _diag([-2.99, -1.99, 0.02, 2.01, 3.01])),
    # :280
    ("OneGroupOfVerticesTwoGroupsOfNodes", ALL_NODES, _diag([-3, -2, 0.01, 2, 3]),
     _diag([-2.99, -1.99, 0.02, 2.01, 3.01])),  # :329
    ("TwoGroupsOfVerticesTwoGroupsOfNodes", ALL_NODES, _diag([-3, -2, 0.01, 2, 3, 12, 11, 10, 10.5, 11.5]),
     _diag([-2.99, -1.99, 0.02, 2.01, 3.01, 11.99, 10.99, 9.99, 10.51, 11.49])),  # :378
]

# multi-solve cases (:454, :530, :632) share these
WARP_SRC = _diag([-3, -2, 0.04, 2, 3])
WARP_T1 = _diag([-2.99, -1.99, 0.05, 2.01, 3.01])
WARP_T2 = _diag([-2.98, -1.98, 0.06, 2.02, 3.02])
WARP_T3 = _diag([-2.96, -1.96, 0.09, 2.04, 3.05])
