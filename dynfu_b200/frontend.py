"""Front end of the frame loop: depth -> points/normals, valid-pixel compaction, live<->canonical correspondences.

    compute_points_normals  cuda::computePointNormals (src/kfusion/imgproc.cpp:27-36, cuda/imgproc.cu:187-226)
    compact_points          the host-side collection of the downloaded cloud (src/dynfu/dyn_fusion.cpp:120-134)
    PointIndex              the per-frame nanoflann KD-tree of DynFusion::findCorrespondingFrame
    find_corresponding      DynFusion::findCorrespondingFrame (src/dynfu/dyn_fusion.cpp:212-242)
    voxel_grid_filter       pcl::VoxelGrid as used by Warpfield::update (src/dynfu/warp_field.cpp:68-72)
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, dptr, dptr2d, farr, lib, stream_ptr


def compute_points_normals(depth, intr, points=None, normals=None):
    """depth uint16/int16 mm [rows, cols] (CUDA) -> (points, normals), float32 [rows, cols, 4], NaN where invalid."""
    rows, cols = depth.shape
    if depth.dtype not in (torch.uint16, torch.int16):
        raise _lib.DfuError(1, "depth must be a 16-bit tensor")
    if points is None:
        points = torch.empty((rows, cols, 4), dtype=torch.float32, device=depth.device)
    if normals is None:
        normals = torch.empty((rows, cols, 4), dtype=torch.float32, device=depth.device)
    check(lib.dfu_compute_points_normals(dptr2d(depth), depth.stride(0) * 2, rows, cols, farr(intr), dptr2d(points),
                                         points.stride(0) * 4, dptr2d(normals), normals.stride(0) * 4, stream_ptr(device=depth.device)))
    return points, normals


def compact_points(points, normals=None, xform=None, capacity=None, sync=True):
    """Valid pixels of a points (and normals) image, raster order, as packed [n, 3] tensors.

    xform: optional 4x4 (or 3x4) rigid transform applied on the way out.  With sync=False the full-capacity
    buffers and the device-side count are returned without a host synchronisation."""
    rows, cols = points.shape[:2]
    cap = rows * cols if capacity is None else int(capacity)
    dev = points.device
    out_v = torch.empty((cap, 3), dtype=torch.float32, device=dev)
    out_n = torch.empty((cap, 3), dtype=torch.float32, device=dev) if normals is not None else None
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    xf = None
    if xform is not None:
        m = np.asarray(xform.cpu() if torch.is_tensor(xform) else xform, dtype=np.float64)
        xf = farr(list(m[:3, :3].reshape(-1)) + list(m[:3, 3]))
    check(lib.dfu_compact_points(dptr2d(points), points.stride(0) * 4, dptr2d(normals) if normals is not None else None,
                                 normals.stride(0) * 4 if normals is not None else 0, rows, cols, xf, dptr(out_v),
                                 dptr(out_n) if out_n is not None else None, cap, dptr(count), stream_ptr(device=points.device)))
    if not sync:
        return out_v, out_n, count
    n = min(int(count.item()), cap)
    return out_v[:n], (out_n[:n] if out_n is not None else None)


def voxel_grid_filter(points, leaf=0.05):
    """pcl::VoxelGrid<PointXYZ> with a cubic leaf (src/dynfu/warp_field.cpp:68-72): centroids, ascending cell index."""
    p = points.reshape(-1, 3).contiguous()
    out = torch.empty_like(p)
    m = C.c_int()
    check(lib.dfu_voxel_grid_filter(dptr(p), p.shape[0], float(leaf), dptr(out), C.byref(m), stream_ptr(device=p.device)))
    return out[:m.value]


class PointIndex:
    """Exact nearest-neighbour index over a point set (uniform grid on the device)."""

    def __init__(self, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._h = C.c_void_p()
        check(lib.dfu_pointindex_create(C.byref(self._h), self.device.index or 0))
        self._pts = None

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.dfu_pointindex_destroy(self._h)
            self._h = None

    def build(self, pts):
        pts = pts.reshape(-1, 3).contiguous()
        self._pts = pts
        check(lib.dfu_pointindex_build(self._h, dptr(pts), pts.shape[0], stream_ptr(device=self.device)))
        return self

    def nearest(self, queries, return_dist=False):
        q = queries.reshape(-1, 3).contiguous()
        idx = torch.empty(q.shape[0], dtype=torch.int32, device=q.device)
        d2 = torch.empty(q.shape[0], dtype=torch.float32, device=q.device) if return_dist else None
        check(lib.dfu_pointindex_nearest(self._h, dptr(q), q.shape[0], dptr(idx), dptr(d2) if d2 is not None else None,
                                         stream_ptr(device=self.device)))
        return (idx, d2) if return_dist else idx

    def find_corresponding(self, canon_v, canon_n, live_v, return_index=False):
        cv = canon_v.reshape(-1, 3).contiguous()
        cn = canon_n.reshape(-1, 3).contiguous() if canon_n is not None else None
        lv = live_v.reshape(-1, 3).contiguous()
        out_v = torch.empty_like(lv)
        out_n = torch.empty_like(lv) if cn is not None else None
        idx = torch.empty(lv.shape[0], dtype=torch.int32, device=lv.device) if return_index else None
        self._pts = cv
        check(lib.dfu_find_corresponding(self._h, dptr(cv), dptr(cn) if cn is not None else None, cv.shape[0], dptr(lv),
                                         lv.shape[0], dptr(out_v), dptr(out_n) if out_n is not None else None,
                                         dptr(idx) if idx is not None else None, stream_ptr(device=self.device)))
        return (out_v, out_n, idx) if return_index else (out_v, out_n)


def find_corresponding(canon_v, canon_n, live_v, index=None):
    """DynFusion::findCorrespondingFrame: the canonical vertex (and normal) nearest to every live vertex."""
    index = index or PointIndex(canon_v.device)
    return index.find_corresponding(canon_v, canon_n, live_v)
