"""Warpfield -- host mirror of the reference's class (include/dynfu/warp_field.hpp:32-78) over the C-ABI."""
import ctypes as C

import torch

from . import _lib
from ._lib import BLEND_REF_COMPOSE, NORMAL_REF, check, dptr, lib, stream_ptr

KNN = 8  # include/dynfu/warp_field.hpp:27


def _f32(t, shape, device):
    t = torch.as_tensor(t, dtype=torch.float32)
    if t.device != device:
        t = t.to(device)
    return t.reshape(shape).contiguous()


class Warpfield:
    """Deformation nodes (dg_v, dg_se3, dg_w) + exact 8-NN + blending, resident on one GPU.

    Nodes are passed as arrays instead of vector<shared_ptr<Node>>: positions [N,3], dual quaternions [N,8]
    (real wxyz, dual wxyz), radial basis weights [N]."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise _lib.DfuError(2, "no CUDA device: dynfu_b200 has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        h = C.c_void_p()
        check(lib.dfu_warpfield_create(C.byref(h), self.device.index or 0))
        self._h = h
        self.epsilon = None

    def _free_pcache(self):
        if getattr(self, "_pcache", None) is not None and lib is not None:
            lib.dfu_pointcache_destroy(self._pcache)
            self._pcache = None

    def __del__(self):
        self._free_pcache()
        h = getattr(self, "_h", None)
        if h:
            lib.dfu_warpfield_destroy(h)
            self._h = None

    @property
    def handle(self):
        return self._h

    # Warpfield::init (src/dynfu/warp_field.cpp:10-28)
    def init(self, epsilon, positions, transformations, radial_basis_weights):
        n = int(torch.as_tensor(positions).reshape(-1, 3).shape[0])
        pos = _f32(positions, (n, 3), self.device)
        dq = _f32(transformations, (n, 8), self.device)
        w = _f32(radial_basis_weights, (n,), self.device)
        check(lib.dfu_warpfield_init(self._h, float(epsilon), dptr(pos), dptr(dq), dptr(w), n, stream_ptr(device=self.device)))
        self.epsilon = float(epsilon)

    def numNodes(self):
        n = C.c_int()
        check(lib.dfu_warpfield_num_nodes(self._h, C.byref(n)))
        return n.value

    # Warpfield::getNodes (src/dynfu/warp_field.cpp:32) -> (positions, transformations, weights) on the device
    def getNodes(self):
        n = self.numNodes()
        pos = torch.empty((n, 3), dtype=torch.float32, device=self.device)
        dq = torch.empty((n, 8), dtype=torch.float32, device=self.device)
        w = torch.empty((n,), dtype=torch.float32, device=self.device)
        check(lib.dfu_warpfield_get_nodes(self._h, dptr(pos), dptr(dq), dptr(w), stream_ptr(device=self.device)))
        return pos, dq, w

    def setTransformations(self, transformations):
        dq = _f32(transformations, (self.numNodes(), 8), self.device)
        check(lib.dfu_warpfield_set_transforms(self._h, dptr(dq), stream_ptr(device=self.device)))

    # Node::updateTransformation(DQ(0,0,0,t)) for every node (src/dynfu/utils/node.cpp:19-23)
    def updateTranslations(self, translations):
        t = _f32(translations, (self.numNodes(), 3), self.device)
        check(lib.dfu_warpfield_update_translations(self._h, dptr(t), stream_ptr(device=self.device)))

    # Warpfield::getUnsupportedVertices (src/dynfu/warp_field.cpp:34-62): the unsupported vertices, in input order
    def getUnsupportedVertices(self, vertices, return_mask=False):
        v = _f32(vertices, (-1, 3), self.device)
        flags = torch.empty(v.shape[0], dtype=torch.uint8, device=self.device)
        check(lib.dfu_warpfield_unsupported(self._h, dptr(v), v.shape[0], dptr(flags), stream_ptr(device=self.device)))
        mask = flags.bool()
        return mask if return_mask else v[mask]

    # Warpfield::update (src/dynfu/warp_field.cpp:64-95): returns (number of unsupported vertices, nodes added)
    def update(self, vertices, blend_mode=BLEND_REF_COMPOSE):
        v = _f32(vertices, (-1, 3), self.device)
        nu, nn = C.c_int(), C.c_int()
        check(lib.dfu_warpfield_update(self._h, dptr(v), v.shape[0], blend_mode, C.byref(nu), C.byref(nn), stream_ptr(device=self.device)))
        return nu.value, nn.value

    # warpToLive for a fixed point set (the canonical frame): neighbours + weights cached across calls
    def warpToLiveCached(self, vertices, normals=None, version=0, blend_mode=BLEND_REF_COMPOSE, normal_mode=NORMAL_REF):
        """`vertices` must be the same tensor (same storage) from call to call; bump `version` when its contents change."""
        v = _f32(vertices, (-1, 3), self.device)
        if getattr(self, "_pcache", None) is None:
            self._pcache = C.c_void_p()
            check(lib.dfu_pointcache_create(C.byref(self._pcache), self.device.index or 0))
        vo = torch.empty_like(v)
        n = no = None
        if normals is not None:
            n = _f32(normals, (-1, 3), self.device)
            no = torch.empty_like(n)
        check(lib.dfu_warpfield_warp_cached(self._h, self._pcache, int(version), dptr(v), dptr(n), v.shape[0], dptr(vo), dptr(no),
                                            blend_mode, normal_mode, stream_ptr(device=self.device)))
        return vo, no

    def cacheStats(self):
        """(bricks the per-voxel neighbour cache holds, bricks currently valid) -- diagnostics"""
        a, b = C.c_longlong(), C.c_longlong()
        check(lib.dfu_warpfield_cache_stats(self._h, C.byref(a), C.byref(b), stream_ptr(device=self.device)))
        return a.value, b.value

    # Warpfield::findNeighborsIndex (src/dynfu/warp_field.cpp:111-122), for [Q,3] vertices at once
    def findNeighborsIndex(self, numNeighbor, vertices, return_dist=False):
        if numNeighbor != KNN:
            raise _lib.DfuError(1, "k is fixed at KNN = 8 (include/dynfu/warp_field.hpp:27)")
        q = _f32(vertices, (-1, 3), self.device)
        idx = torch.empty((q.shape[0], KNN), dtype=torch.int32, device=self.device)
        d2 = torch.empty((q.shape[0], KNN), dtype=torch.float32, device=self.device) if return_dist else None
        check(lib.dfu_warpfield_knn(self._h, dptr(q), q.shape[0], dptr(idx), dptr(d2), stream_ptr(device=self.device)))
        return (idx, d2) if return_dist else idx

    # Warpfield::calcDQB (src/dynfu/warp_field.cpp:127-148)
    def calcDQB(self, points, blend_mode=BLEND_REF_COMPOSE):
        p = _f32(points, (-1, 3), self.device)
        out = torch.empty((p.shape[0], 8), dtype=torch.float32, device=self.device)
        check(lib.dfu_warpfield_blend(self._h, dptr(p), p.shape[0], dptr(out), blend_mode, stream_ptr(device=self.device)))
        return out

    # Warpfield::warpToLive (src/dynfu/warp_field.cpp:150-171): returns (vertices, normals)
    def warpToLive(self, vertices, normals=None, blend_mode=BLEND_REF_COMPOSE, normal_mode=NORMAL_REF):
        v = _f32(vertices, (-1, 3), self.device)
        vo = torch.empty_like(v)
        n = no = None
        if normals is not None:
            n = _f32(normals, (-1, 3), self.device)
            no = torch.empty_like(n)
        check(lib.dfu_warpfield_warp(self._h, dptr(v), dptr(n), v.shape[0], dptr(vo), dptr(no), blend_mode, normal_mode,
                                     stream_ptr(device=self.device)))
        return vo, no
