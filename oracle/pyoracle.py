"""ctypes/numpy binding of the CPU oracle (oracle/libdynfu_oracle.so).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by anything under dynfu_b200/ (the product).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

BLEND_REF_COMPOSE = 0
BLEND_DQB_SUM = 1
NORMAL_REF = 0
NORMAL_ROTATE_ONLY = 1


class SolverParams(C.Structure):
    _fields_ = [
        ("num_iter", C.c_int),
        ("nonlinear_iter", C.c_int),
        ("linear_iter", C.c_int),
        ("tukey_offset", C.c_float),
        ("psi_data", C.c_float),
        ("lambda_", C.c_float),
        ("psi_reg", C.c_float),
        ("pcg_tol", C.c_double),
        ("early_out", C.c_int),
        ("reg_mode", C.c_int),
    ]


def build(verbose=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    r = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)


_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_u16p = C.POINTER(C.c_uint16)
_u32p = C.POINTER(C.c_uint32)


def _f(a):
    return a.ctypes.data_as(_fp) if a is not None else None


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


class Oracle:
    """One loaded oracle library.  kind='brute' (default checker) or 'nanoflann' (oracle/_ref)."""

    def __init__(self, kind="brute"):
        path = os.path.join(_HERE, "libdynfu_oracle.so") if kind == "brute" else os.path.join(
            _HERE, "_ref", "libdynfu_oracle_nf.so")
        if not os.path.exists(path) and kind == "brute":
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.kind = kind
        self.path = path
        L = self.lib = C.CDLL(path)
        L.orc_dq_from_euler.argtypes = [C.c_float] * 6 + [_fp]
        L.orc_dq_from_rot_trans.argtypes = [_fp, _fp, _fp]
        L.orc_dq_from_rodrigues.argtypes = [_fp, _fp, _fp]
        for n in ("add", "sub", "mul"):
            getattr(L, "orc_dq_" + n).argtypes = [_fp, _fp, _fp]
        L.orc_dq_scale.argtypes = [_fp, C.c_float, _fp]
        L.orc_dq_conj.argtypes = [_fp, _fp]
        L.orc_dq_normalize.argtypes = [_fp, _fp]
        L.orc_dq_normalize.restype = C.c_int
        L.orc_dq_get_translation.argtypes = [_fp, _fp]
        for n in ("roll", "pitch", "yaw"):
            fn = getattr(L, "orc_dq_get_" + n)
            fn.argtypes = [_fp]
            fn.restype = C.c_float
        L.orc_dq_get_rodrigues.argtypes = [_fp, _fp]
        L.orc_dq_transform_vertex.argtypes = [_fp, _fp, _fp]
        L.orc_dq_transform_normal.argtypes = [_fp, _fp, C.c_int, _fp]
        L.orc_dq_to_string.argtypes = [_fp, C.c_char_p, C.c_int]
        L.orc_dq_to_string.restype = C.c_int
        L.orc_node_weight.argtypes = [_fp, C.c_float, _fp]
        L.orc_node_weight.restype = C.c_float
        L.orc_knn.argtypes = [_fp, C.c_int, _fp, C.c_long, C.c_int, _ip, _fp]
        L.orc_knn.restype = C.c_long
        L.orc_blend.argtypes = [_fp, _fp, _fp, C.c_int, _fp, C.c_long, C.c_int, _fp]
        L.orc_warp.argtypes = [_fp, _fp, _fp, C.c_int, _fp, _fp, C.c_long, C.c_int, C.c_int, _fp, _fp]
        L.orc_compute_dists.argtypes = [_u16p, C.c_size_t, _u16p, C.c_size_t, C.c_int, C.c_int, _fp]
        L.orc_points_normals.argtypes = [_u16p, C.c_size_t, C.c_int, C.c_int, _fp, _fp, _fp]
        L.orc_compact_points.argtypes = [_fp, _fp, C.c_int, C.c_int, _fp, _fp, _fp]
        L.orc_compact_points.restype = C.c_long
        L.orc_find_corresponding.argtypes = [_fp, _fp, C.c_int, _fp, C.c_long, _fp, _fp, _ip]
        L.orc_find_corresponding.restype = C.c_long
        L.orc_unsupported.argtypes = [_fp, _fp, C.c_int, _fp, C.c_long, C.POINTER(C.c_uint8)]
        L.orc_unsupported.restype = C.c_long
        L.orc_voxel_grid.argtypes = [_fp, C.c_long, _fp, _fp, C.c_int]
        L.orc_voxel_grid.restype = C.c_long
        L.orc_warpfield_update.argtypes = [_fp, _fp, _fp, C.c_int, C.c_float, _fp, C.c_long, C.c_int, _fp, _fp, _fp]
        L.orc_warpfield_update.restype = C.c_long
        L.orc_raycast.argtypes = [_u32p, _ip, _fp, C.c_float, _fp, _fp, _fp, C.c_int, C.c_int, C.c_float, C.c_float, _fp, _fp, _u16p]
        _dp = C.POINTER(C.c_double)
        L.orc_solve_p2plane.argtypes = [_fp, _fp, _fp, C.c_int, _fp, _fp, _fp, C.c_long, C.POINTER(SolverParams), _dp, _dp]
        L.orc_solve_p2plane.restype = C.c_int
        L.orc_energy_p2plane.argtypes = [_fp, _fp, C.c_int, _fp, _fp, _fp, C.c_long, C.POINTER(SolverParams), _dp, _dp]
        L.orc_energy_p2plane.restype = C.c_double
        L.orc_marching_cubes.argtypes = [_u32p, _ip, _fp, C.POINTER(C.c_int8), _fp, _ip, C.c_long]
        L.orc_marching_cubes.restype = C.c_long
        L.orc_float2half.argtypes = [C.c_float]
        L.orc_float2half.restype = C.c_uint16
        L.orc_half2float.argtypes = [C.c_uint16]
        L.orc_half2float.restype = C.c_float
        L.orc_tsdf_clear.argtypes = [_u32p, _ip, C.c_int, C.c_int]
        L.orc_tsdf_integrate.argtypes = [_u32p, _ip, _fp, C.c_float, C.c_int, _fp, _fp, _u16p, C.c_size_t, C.c_int,
                                         C.c_int, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp]
        L.orc_tsdf_integrate.restype = C.c_long
        L.orc_trunc_dist.argtypes = [C.c_float, _fp]
        L.orc_trunc_dist.restype = C.c_float
        L.orc_tukey.argtypes = [C.c_float, C.c_float, _fp]
        L.orc_tukey.restype = C.c_float
        L.orc_huber.argtypes = [C.c_float, C.c_float]
        L.orc_huber.restype = C.c_float
        L.orc_huber_weights.argtypes = [_fp, _fp, C.c_int, C.c_float, _fp]
        L.orc_solve.argtypes = [_fp, _fp, _fp, C.c_int, _fp, _fp, C.c_long, C.POINTER(SolverParams), _dp, _dp]
        L.orc_solve.restype = C.c_int
        L.orc_energy.argtypes = [_fp, _fp, C.c_int, _fp, _fp, C.c_long, C.POINTER(SolverParams), _dp, _dp]
        L.orc_energy.restype = C.c_double
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]

    # ---- dual quaternions ------------------------------------------------------------------
    def dq_from_euler(self, yaw, pitch, roll, x, y, z):
        out = np.zeros(8, np.float32)
        self.lib.orc_dq_from_euler(yaw, pitch, roll, x, y, z, _f(out))
        return out

    def dq_from_rot_trans(self, rot, t):
        out = np.zeros(8, np.float32)
        rot, t = _f32(rot), _f32(t)
        self.lib.orc_dq_from_rot_trans(_f(rot), _f(t), _f(out))
        return out

    def dq_from_rodrigues(self, rod, t):
        out = np.zeros(8, np.float32)
        rod, t = _f32(rod), _f32(t)
        self.lib.orc_dq_from_rodrigues(_f(rod), _f(t), _f(out))
        return out

    def _bin(self, name, a, b):
        out = np.zeros(8, np.float32)
        a, b = _f32(a), _f32(b)
        getattr(self.lib, "orc_dq_" + name)(_f(a), _f(b), _f(out))
        return out

    def dq_add(self, a, b):
        return self._bin("add", a, b)

    def dq_sub(self, a, b):
        return self._bin("sub", a, b)

    def dq_mul(self, a, b):
        return self._bin("mul", a, b)

    def dq_scale(self, a, s):
        out = np.zeros(8, np.float32)
        a = _f32(a)
        self.lib.orc_dq_scale(_f(a), s, _f(out))
        return out

    def dq_conj(self, a):
        out = np.zeros(8, np.float32)
        a = _f32(a)
        self.lib.orc_dq_conj(_f(a), _f(out))
        return out

    def dq_normalize(self, a):
        out = np.zeros(8, np.float32)
        a = _f32(a)
        rc = self.lib.orc_dq_normalize(_f(a), _f(out))
        if rc != 0:
            raise AssertionError("magnitude > epsilon")  # dual_quaternion.hpp:141
        return out

    def dq_translation(self, a):
        out = np.zeros(3, np.float32)
        a = _f32(a)
        self.lib.orc_dq_get_translation(_f(a), _f(out))
        return out

    def dq_roll(self, a):
        a = _f32(a)
        return self.lib.orc_dq_get_roll(_f(a))

    def dq_pitch(self, a):
        a = _f32(a)
        return self.lib.orc_dq_get_pitch(_f(a))

    def dq_yaw(self, a):
        a = _f32(a)
        return self.lib.orc_dq_get_yaw(_f(a))

    def dq_euler_angles(self, a):
        return np.array([self.dq_roll(a), self.dq_pitch(a), self.dq_yaw(a)], np.float32)

    def dq_rodrigues(self, a):
        out = np.zeros(3, np.float32)
        a = _f32(a)
        self.lib.orc_dq_get_rodrigues(_f(a), _f(out))
        return out

    def dq_transform_vertex(self, a, v):
        out = np.zeros(3, np.float32)
        a, v = _f32(a), _f32(v)
        self.lib.orc_dq_transform_vertex(_f(a), _f(v), _f(out))
        return out

    def dq_transform_normal(self, a, n, normal_mode=NORMAL_REF):
        out = np.zeros(3, np.float32)
        a, n = _f32(a), _f32(n)
        self.lib.orc_dq_transform_normal(_f(a), _f(n), normal_mode, _f(out))
        return out

    def dq_to_string(self, a):
        buf = C.create_string_buffer(256)
        a = _f32(a)
        self.lib.orc_dq_to_string(_f(a), buf, 256)
        return buf.value.decode()

    # ---- warp field --------------------------------------------------------------------------
    def node_weight(self, node_pos, dg_w, p):
        node_pos, p = _f32(node_pos), _f32(p)
        return self.lib.orc_node_weight(_f(node_pos), dg_w, _f(p))

    def knn(self, nodes, queries, k=8, return_dist=False):
        nodes = _f32(nodes, (-1, 3))
        queries = _f32(queries, (-1, 3))
        Q = queries.shape[0]
        idx = np.empty((Q, k), np.int32)
        d2 = np.empty((Q, k), np.float32)
        ties = self.lib.orc_knn(_f(nodes), nodes.shape[0], _f(queries), Q, k, idx.ctypes.data_as(_ip), _f(d2))
        if return_dist:
            return idx, d2, ties
        return idx, ties

    def blend(self, pos, dq, dg_w, pts, mode=BLEND_REF_COMPOSE):
        pos, dq, dg_w, pts = _f32(pos, (-1, 3)), _f32(dq, (-1, 8)), _f32(dg_w), _f32(pts, (-1, 3))
        out = np.empty((pts.shape[0], 8), np.float32)
        self.lib.orc_blend(_f(pos), _f(dq), _f(dg_w), pos.shape[0], _f(pts), pts.shape[0], mode, _f(out))
        return out

    def warp(self, pos, dq, dg_w, v, n=None, blend_mode=BLEND_REF_COMPOSE, normal_mode=NORMAL_REF):
        pos, dq, dg_w, v = _f32(pos, (-1, 3)), _f32(dq, (-1, 8)), _f32(dg_w), _f32(v, (-1, 3))
        vo = np.empty_like(v)
        no = None
        if n is not None:
            n = _f32(n, (-1, 3))
            no = np.empty_like(n)
        self.lib.orc_warp(_f(pos), _f(dq), _f(dg_w), pos.shape[0], _f(v), _f(n), v.shape[0], blend_mode, normal_mode,
                          _f(vo), _f(no))
        return (vo, no) if n is not None else vo

    # ---- TSDF --------------------------------------------------------------------------------
    def compute_dists(self, depth, intr):
        depth = np.ascontiguousarray(depth, np.uint16)
        rows, cols = depth.shape
        out = np.empty_like(depth)
        intr = _f32(intr)
        self.lib.orc_compute_dists(depth.ctypes.data_as(_u16p), cols * 2, out.ctypes.data_as(_u16p), cols * 2, rows,
                                   cols, _f(intr))
        return out

    def points_normals(self, depth, intr):
        """cuda::computePointNormals: (rows, cols, 4) points and normals, NaN where invalid."""
        depth = np.ascontiguousarray(depth, np.uint16)
        rows, cols = depth.shape
        pts = np.empty((rows, cols, 4), np.float32)
        nrm = np.empty((rows, cols, 4), np.float32)
        self.lib.orc_points_normals(depth.ctypes.data_as(_u16p), cols * 2, rows, cols, _f(_f32(intr)), _f(pts), _f(nrm))
        return pts, nrm

    def compact_points(self, pts4, nrm4=None, xform=None):
        pts4 = _f32(pts4)
        rows, cols = pts4.shape[:2]
        nrm4 = _f32(nrm4) if nrm4 is not None else None
        xf = _f32(xform) if xform is not None else None
        v = np.empty((rows * cols, 3), np.float32)
        n = np.empty((rows * cols, 3), np.float32) if nrm4 is not None else None
        cnt = self.lib.orc_compact_points(_f(pts4), _f(nrm4), rows, cols, _f(xf), _f(v), _f(n))
        return (v[:cnt].copy(), n[:cnt].copy()) if n is not None else v[:cnt].copy()

    def find_corresponding(self, canon_v, canon_n, live_v):
        """DynFusion::findCorrespondingFrame: (vertices, normals, indices, ties)."""
        canon_v, live_v = _f32(canon_v, (-1, 3)), _f32(live_v, (-1, 3))
        canon_n = _f32(canon_n, (-1, 3)) if canon_n is not None else None
        P = live_v.shape[0]
        ov = np.empty((P, 3), np.float32)
        on = np.empty((P, 3), np.float32) if canon_n is not None else None
        idx = np.empty(P, np.int32)
        ties = self.lib.orc_find_corresponding(_f(canon_v), _f(canon_n), canon_v.shape[0], _f(live_v), P, _f(ov), _f(on),
                                               idx.ctypes.data_as(_ip))
        return ov, on, idx, ties

    def unsupported(self, pos, dg_w, verts):
        """Warpfield::getUnsupportedVertices: boolean mask over the vertices."""
        pos, dg_w, verts = _f32(pos, (-1, 3)), _f32(dg_w), _f32(verts, (-1, 3))
        flags = np.zeros(max(verts.shape[0], 1), np.uint8)
        self.lib.orc_unsupported(_f(pos), _f(dg_w), pos.shape[0], _f(verts), verts.shape[0],
                                 flags.ctypes.data_as(C.POINTER(C.c_uint8)))
        return flags[:verts.shape[0]].astype(bool)

    def voxel_grid(self, pts, leaf=0.05, order_mode=0):
        """pcl::VoxelGrid centroids (None if PCL would refuse the grid)."""
        pts = _f32(pts, (-1, 3))
        out = np.empty((max(pts.shape[0], 1), 3), np.float32)
        lf = np.full(3, leaf, np.float32)
        m = self.lib.orc_voxel_grid(_f(pts), pts.shape[0], _f(lf), _f(out), order_mode)
        return None if m < 0 else out[:m].copy()

    def warpfield_update(self, pos, dq, dg_w, epsilon, verts, blend_mode=BLEND_REF_COMPOSE):
        pos, dq, dg_w, verts = _f32(pos, (-1, 3)), _f32(dq, (-1, 8)), _f32(dg_w), _f32(verts, (-1, 3))
        N, P = pos.shape[0], verts.shape[0]
        po, qo, wo = np.empty((N + P, 3), np.float32), np.empty((N + P, 8), np.float32), np.empty(N + P, np.float32)
        n = self.lib.orc_warpfield_update(_f(pos), _f(dq), _f(dg_w), N, epsilon, _f(verts), P, blend_mode, _f(po), _f(qo),
                                          _f(wo))
        return po[:n].copy(), qo[:n].copy(), wo[:n].copy()

    def raycast(self, vol, voxel, trunc, cam2vol, rinv, intr, rows, cols, step_factor=0.75, grad_factor=0.75, want_depth=False):
        """TsdfVolume::raycast: (points [rows, cols, 4], normals [rows, cols, 4][, depth u16])."""
        vol = np.ascontiguousarray(vol, np.uint32)
        dims = np.array(vol.shape[::-1], np.int32)
        pts = np.empty((rows, cols, 4), np.float32)
        nrm = np.empty((rows, cols, 4), np.float32)
        dep = np.zeros((rows, cols), np.uint16) if want_depth else None
        self.lib.orc_raycast(vol.ctypes.data_as(_u32p), dims.ctypes.data_as(_ip), _f(_f32(voxel)), float(trunc), _f(_f32(cam2vol)),
                             _f(_f32(rinv)), _f(_f32(intr)), rows, cols, float(step_factor), float(grad_factor), _f(pts), _f(nrm),
                             dep.ctypes.data_as(_u16p) if dep is not None else None)
        return (pts, nrm, dep) if want_depth else (pts, nrm)

    def marching_cubes(self, vol, volume_size, tri_table, capacity=None):
        """MarchingCubes::run: (vertices [n, 4], cube ids [n]) of the zero level set of vol (uint32 [z][y][x])."""
        vol = np.ascontiguousarray(vol, np.uint32)
        dims = np.array(vol.shape[::-1], np.int32)
        tri = np.ascontiguousarray(tri_table, np.int8).reshape(256, 16)
        cap = int(capacity) if capacity is not None else 15 * int(vol.size // 8 + 1024)
        verts = np.empty((cap, 4), np.float32)
        ids = np.empty(cap, np.int32)
        n = self.lib.orc_marching_cubes(vol.ctypes.data_as(_u32p), dims.ctypes.data_as(_ip), _f(_f32(volume_size)),
                                        tri.ctypes.data_as(C.POINTER(C.c_int8)), _f(verts), ids.ctypes.data_as(_ip), cap)
        m = min(int(n), cap)
        return verts[:m].copy(), ids[:m].copy(), int(n)

    def float2half(self, f):
        return self.lib.orc_float2half(f)

    def half2float(self, h):
        return self.lib.orc_half2float(h)

    def trunc_dist(self, requested, voxel):
        voxel = _f32(voxel)
        return self.lib.orc_trunc_dist(requested, _f(voxel))

    def tsdf_clear(self, vol, z0=0, z1=None):
        dims = np.array(vol.shape[::-1], np.int32)  # vol is [z][y][x]
        z1 = dims[2] if z1 is None else z1
        self.lib.orc_tsdf_clear(vol.ctypes.data_as(_u32p), dims.ctypes.data_as(_ip), z0, z1)

    def tsdf_integrate(self, vol, voxel, trunc, max_weight, vol2cam, intr, dists, nodes=None,
                       blend_mode=BLEND_REF_COMPOSE, z0=0, z1=None, f32_out=None):
        """vol: uint32 array [z][y][x] (ushort2 {half tsdf, u16 weight} packed little endian), in place.
        vol2cam: 12 floats (row-major R then t).  nodes: None or (pos[N,3], dq[N,8], dg_w[N])."""
        assert vol.dtype == np.uint32 and vol.flags.c_contiguous
        dims = np.array(vol.shape[::-1], np.int32)
        z1 = int(dims[2]) if z1 is None else z1
        voxel, vol2cam, intr = _f32(voxel), _f32(vol2cam), _f32(intr)
        dists = np.ascontiguousarray(dists, np.uint16)
        rows, cols = dists.shape
        if nodes is not None:
            pos, dq, dg_w = _f32(nodes[0], (-1, 3)), _f32(nodes[1], (-1, 8)), _f32(nodes[2])
            N = pos.shape[0]
        else:
            pos = dq = dg_w = None
            N = 0
        return self.lib.orc_tsdf_integrate(vol.ctypes.data_as(_u32p), dims.ctypes.data_as(_ip), _f(voxel), trunc,
                                           max_weight, _f(vol2cam), _f(intr), dists.ctypes.data_as(_u16p), cols * 2,
                                           rows, cols, _f(pos), _f(dq), _f(dg_w), N, blend_mode, z0, z1, _f(f32_out))

    # ---- solver ------------------------------------------------------------------------------
    def tukey(self, tukey_offset, c, err):
        err = _f32(err)
        return self.lib.orc_tukey(tukey_offset, c, _f(err))

    def huber(self, k, e):
        return self.lib.orc_huber(k, e)

    def huber_weights(self, pos, dq, psi_reg):
        pos, dq = _f32(pos, (-1, 3)), _f32(dq, (-1, 8))
        out = np.empty(pos.shape[0], np.float32)
        self.lib.orc_huber_weights(_f(pos), _f(dq), pos.shape[0], psi_reg, _f(out))
        return out

    def solve(self, pos, dq, dg_w, canon, live, params):
        """Returns (t[N,3] float64, dq_new[N,8] float32, stats[4])."""
        pos, dq, dg_w = _f32(pos, (-1, 3)), _f32(dq, (-1, 8)).copy(), _f32(dg_w)
        canon, live = _f32(canon, (-1, 3)), _f32(live, (-1, 3))
        N = pos.shape[0]
        t = np.zeros((N, 3), np.float64)
        stats = np.zeros(4, np.float64)
        rc = self.lib.orc_solve(_f(pos), _f(dq), _f(dg_w), N, _f(canon), _f(live), canon.shape[0], C.byref(params),
                                t.ctypes.data_as(_dp), stats.ctypes.data_as(_dp))
        if rc != 0:
            raise ValueError("orc_solve: precondition N >= 8 violated")
        return t, dq, stats

    def energy(self, pos, dg_w, canon, live, params, t, t_tukey=None):
        pos, dg_w = _f32(pos, (-1, 3)), _f32(dg_w)
        canon, live = _f32(canon, (-1, 3)), _f32(live, (-1, 3))
        t = np.ascontiguousarray(t, np.float64)
        tt = t if t_tukey is None else np.ascontiguousarray(t_tukey, np.float64)
        return self.lib.orc_energy(_f(pos), _f(dg_w), pos.shape[0], _f(canon), _f(live), canon.shape[0],
                                   C.byref(params), t.ctypes.data_as(_dp), tt.ctypes.data_as(_dp))

    # ---- north-star extension: point-to-plane data term, rigid increment per node (no reference implementation) ----
    def solve_p2plane(self, pos, dq, dg_w, canon, live, live_n, params):
        """returns (X [N, 12] increments (R row-major, t), dq_out [N, 8], stats [E0, E, pcg, gn])"""
        pos, dq, dg_w = _f32(pos, (-1, 3)), _f32(dq, (-1, 8)).copy(), _f32(dg_w)
        canon, live, live_n = _f32(canon, (-1, 3)), _f32(live, (-1, 3)), _f32(live_n, (-1, 3))
        N, P = pos.shape[0], canon.shape[0]
        X = np.zeros((N, 12), np.float64)
        stats = np.zeros(4, np.float64)
        _dp = C.POINTER(C.c_double)
        rc = self.lib.orc_solve_p2plane(_f(pos), _f(dq), _f(dg_w), N, _f(canon), _f(live), _f(live_n), P, C.byref(params),
                                        X.ctypes.data_as(_dp), stats.ctypes.data_as(_dp))
        if rc != 0:
            raise ValueError("orc_solve_p2plane failed: %d" % rc)
        return X, dq, stats

    def energy_p2plane(self, pos, dg_w, canon, live, live_n, params, X, X_tukey=None):
        pos, dg_w = _f32(pos, (-1, 3)), _f32(dg_w)
        canon, live, live_n = _f32(canon, (-1, 3)), _f32(live, (-1, 3)), _f32(live_n, (-1, 3))
        X = np.ascontiguousarray(X, np.float64).reshape(-1, 12)
        Xt = X if X_tukey is None else np.ascontiguousarray(X_tukey, np.float64).reshape(-1, 12)
        _dp = C.POINTER(C.c_double)
        return self.lib.orc_energy_p2plane(_f(pos), _f(dg_w), pos.shape[0], _f(canon), _f(live), _f(live_n), canon.shape[0],
                                           C.byref(params), X.ctypes.data_as(_dp), Xt.ctypes.data_as(_dp))

    def set_num_threads(self, n):
        self.lib.orc_set_num_threads(int(n))

    def num_threads(self):
        return self.lib.orc_num_threads()


def default_params(num_iter=32, nonlinear_iter=16, linear_iter=256, tukey_offset=4.652, psi_data=1e-2, lambda_=0.0,
                   psi_reg=1e-4, pcg_tol=1e-12, early_out=1, reg_mode=0):
    """Defaults = test/opt_optimisation_test.cpp:38-44,115-122."""
    return SolverParams(num_iter, nonlinear_iter, linear_iter, tukey_offset, psi_data, lambda_, psi_reg, pcg_tol,
                        early_out, reg_mode)


class RefCuda:
    """oracle/_ref/libdynfu_ref_cuda.so: the reference's OWN CUDA kernels (TSDF integrate / clear / raycast, compute_dists,
    computePointNormals, marching cubes) compiled where they lie behind oracle/ref_shim (recipe: oracle/Makefile, target
    refcuda).  Needs a GPU; arguments are torch CUDA tensors.  TEST INFRASTRUCTURE ONLY (`-m gpu` tests)."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libdynfu_ref_cuda.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        L = self.lib = C.CDLL(path)
        vp = C.c_void_p
        L.ref_clear_volume.argtypes = [vp, _ip]
        L.ref_integrate.argtypes = [vp, _ip, _fp, C.c_float, C.c_int, vp, C.c_size_t, C.c_int, C.c_int, _fp, _fp]
        L.ref_compute_dists.argtypes = [vp, C.c_size_t, vp, C.c_size_t, C.c_int, C.c_int, _fp]
        L.ref_points_normals.argtypes = [vp, C.c_size_t, C.c_int, C.c_int, _fp, vp, vp]
        L.ref_raycast_points.argtypes = [vp, _ip, _fp, C.c_float, C.c_int, _fp, _fp, _fp, C.c_int, C.c_int, vp, vp, C.c_float,
                                         C.c_float]
        L.ref_mc_tables.argtypes = [_ip, _ip, _ip]
        L.ref_mc_tables.restype = None
        L.ref_marching_cubes.argtypes = [vp, _fp, C.c_float, C.c_int, _fp, vp, C.c_int, vp, C.c_int, _ip, _ip]

    @staticmethod
    def _dims(vol):
        return np.array(vol.shape[::-1], np.int32)

    @staticmethod
    def _ok(rc, what):
        if rc != 0:
            raise RuntimeError("reference kernel %s failed" % what)

    def clear_volume(self, vol):
        """vol: int32 CUDA tensor [z][y][x] (packed ushort2)"""
        d = self._dims(vol)
        self._ok(self.lib.ref_clear_volume(vol.data_ptr(), d.ctypes.data_as(_ip)), "clear_volume")

    def integrate(self, vol, voxel, trunc, max_weight, vol2cam, intr, dists):
        """device::integrate on the int32 CUDA tensor vol [z][y][x]; dists: 16-bit CUDA tensor [rows][cols] of half bits"""
        d = self._dims(vol)
        voxel, vol2cam, intr = _f32(voxel), _f32(vol2cam), _f32(intr)
        rows, cols = dists.shape
        self._ok(self.lib.ref_integrate(vol.data_ptr(), d.ctypes.data_as(_ip), _f(voxel), float(trunc), int(max_weight),
                                        dists.data_ptr(), dists.stride(0) * 2, rows, cols, _f(vol2cam), _f(intr)), "integrate")

    def compute_dists(self, depth, dists_out, intr):
        rows, cols = depth.shape
        intr = _f32(intr)
        self._ok(self.lib.ref_compute_dists(depth.data_ptr(), depth.stride(0) * 2, dists_out.data_ptr(), dists_out.stride(0) * 2,
                                            rows, cols, _f(intr)), "compute_dists")

    def points_normals(self, depth, points_out, normals_out, intr):
        """points_out / normals_out: float32 CUDA tensors [rows][cols][4]"""
        rows, cols = depth.shape
        intr = _f32(intr)
        self._ok(self.lib.ref_points_normals(depth.data_ptr(), depth.stride(0) * 2, rows, cols, _f(intr), points_out.data_ptr(),
                                             normals_out.data_ptr()), "computePointNormals")

    def raycast_points(self, vol, voxel, trunc, max_weight, cam2vol, rinv, intr, points_out, normals_out, step_factor,
                       delta_factor):
        d = self._dims(vol)
        voxel, cam2vol, rinv, intr = _f32(voxel), _f32(cam2vol), _f32(rinv), _f32(intr)
        rows, cols = points_out.shape[:2]
        self._ok(self.lib.ref_raycast_points(vol.data_ptr(), d.ctypes.data_as(_ip), _f(voxel), float(trunc), int(max_weight),
                                             _f(cam2vol), _f(rinv), _f(intr), rows, cols, points_out.data_ptr(),
                                             normals_out.data_ptr(), float(step_factor), float(delta_factor)), "raycast")

    def mc_tables(self):
        """(edgeTable[256], triTable[256][16], numVertsTable[256]) of src/kfusion/marching_cubes.cpp:66-354"""
        e = np.zeros(256, np.int32)
        t = np.zeros(256 * 16, np.int32)
        n = np.zeros(256, np.int32)
        self.lib.ref_mc_tables(e.ctypes.data_as(_ip), t.ctypes.data_as(_ip), n.ctypes.data_as(_ip))
        return e, t.reshape(256, 16), n

    def marching_cubes(self, vol, voxel, trunc, max_weight, volume_size, occupied, triangles):
        """the reference's marching cubes on a 128^3 volume; occupied: int32 CUDA [3][max_voxels] scratch, triangles: float32
        CUDA [max_vertices][4].  Returns (n_voxels, n_vertices); voxel emission order is atomic-dependent."""
        assert tuple(vol.shape) == (128, 128, 128)
        voxel, volume_size = _f32(voxel), _f32(volume_size)
        nv = np.zeros(1, np.int32)
        nt = np.zeros(1, np.int32)
        self._ok(self.lib.ref_marching_cubes(vol.data_ptr(), _f(voxel), float(trunc), int(max_weight), _f(volume_size),
                                             occupied.data_ptr(), occupied.shape[1], triangles.data_ptr(), triangles.shape[0],
                                             nv.ctypes.data_as(_ip), nt.ctypes.data_as(_ip)), "marching_cubes")
        return int(nv[0]), int(nt[0])
