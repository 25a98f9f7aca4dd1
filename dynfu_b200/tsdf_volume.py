"""TsdfVolume -- host mirror of kfusion::cuda::TsdfVolume (include/kfusion/cuda/tsdf_volume.hpp:7-73)."""
import torch

from . import _lib
from ._lib import BLEND_REF_COMPOSE, check, dptr, farr, iarr, lib, stream_ptr


def compute_dists(depth, intr, out=None):
    """cuda::computeDists (src/kfusion/imgproc.cpp:38-41): depth uint16 mm [rows, cols] (CUDA) -> half bits."""
    rows, cols = depth.shape
    if depth.dtype not in (torch.uint16, torch.int16):
        raise _lib.DfuError(1, "depth must be a 16-bit tensor")
    if out is None:
        out = torch.empty_like(depth)
    check(lib.dfu_compute_dists(dptr(depth), depth.stride(0) * 2, dptr(out), out.stride(0) * 2, rows, cols, farr(intr),
                                stream_ptr(device=depth.device)))
    return out


class TsdfVolume:
    """Dense volume of ushort2 {half tsdf, u16 weight}, x fastest (include/kfusion/cuda/device.hpp:20-35,59-67),
    held as an int32 tensor [dz, dy, dx].  z0/z1 select the z-slab this rank owns (multi-GPU sharding)."""

    def __init__(self, dims, device=None, z0=0, z1=None, size=(3.0, 3.0, 3.0)):
        self.dims = tuple(int(d) for d in dims)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.z0 = int(z0)
        self.z1 = self.dims[2] if z1 is None else int(z1)
        self.size = tuple(float(s) for s in size)  # metres (tsdf_volume.cpp:24)
        self.trunc_dist = 0.03  # tsdf_volume.cpp:21
        self.max_weight = 128  # tsdf_volume.cpp:22
        self.pose = torch.eye(4, dtype=torch.float64)  # Affine3f::Identity (tsdf_volume.cpp:25)
        # only the owned slab is allocated; data_full_offset maps slab plane 0 to global plane z0
        self.data = torch.zeros((self.z1 - self.z0, self.dims[1], self.dims[0]), dtype=torch.int32, device=self.device)
        self.setTruncDist(self.trunc_dist)

    # -- getters / setters of tsdf_volume.cpp:40-72
    def getDims(self):
        return self.dims

    def getVoxelSize(self):
        t = torch.tensor(self.size, dtype=torch.float32) / torch.tensor(self.dims, dtype=torch.float32)
        return tuple(float(x) for x in t)

    def getSize(self):
        return self.size

    def setSize(self, size):
        self.size = tuple(float(s) for s in size)
        self.setTruncDist(self.trunc_dist)

    def getTruncDist(self):
        return self.trunc_dist

    def setTruncDist(self, distance):  # tsdf_volume.cpp:57-61
        self.trunc_dist = float(lib.dfu_tsdf_trunc_dist(float(distance), farr(self.getVoxelSize())))

    def getMaxWeight(self):
        return self.max_weight

    def setMaxWeight(self, w):
        self.max_weight = int(w)

    def getPose(self):
        return self.pose

    def setPose(self, pose4x4):
        self.pose = torch.as_tensor(pose4x4, dtype=torch.float64).reshape(4, 4).clone()

    def _base_ptr(self):
        """Pointer the C-ABI expects: address of global plane 0 (the slab may start at z0 > 0)."""
        import ctypes as C
        plane_bytes = self.dims[0] * self.dims[1] * 4
        return C.c_void_p(self.data.data_ptr() - self.z0 * plane_bytes)

    # TsdfVolume::clear (tsdf_volume.cpp:74-80)
    def clear(self):
        check(lib.dfu_tsdf_clear(self._base_ptr(), iarr(self.dims), self.z0, self.z1, stream_ptr(device=self.device)))

    # TsdfVolume::integrate (tsdf_volume.cpp:82-93); warpfield=None is the reference's rigid integrator
    def integrate(self, dists, camera_pose, intr, warpfield=None, blend_mode=BLEND_REF_COMPOSE):
        cam = torch.as_tensor(camera_pose, dtype=torch.float64).reshape(4, 4)
        vol2cam = torch.linalg.inv(cam) @ self.pose  # camera_pose.inv() * pose_
        v2c = [float(x) for x in vol2cam[:3, :3].reshape(-1)] + [float(x) for x in vol2cam[:3, 3]]
        rows, cols = dists.shape
        check(lib.dfu_tsdf_integrate(self._base_ptr(), iarr(self.dims), farr(self.getVoxelSize()), self.trunc_dist,
                                     self.max_weight, farr(v2c), farr(intr), dptr(dists), dists.stride(0) * 2, rows, cols,
                                     warpfield.handle if warpfield is not None else None, blend_mode, self.z0, self.z1,
                                     stream_ptr(device=self.device)))

    # kfusion::cuda::MarchingCubes::run (src/kfusion/marching_cubes.cpp:20-63) on this volume
    def marchingCubes(self, capacity=None, with_cube_ids=False):
        """triangle vertices of the zero level set, float32 [n, 4] = (x, y, z, 1) in volume-local metres (3 per triangle), in
        the library's fixed order; n is read back from the device (one synchronisation, like the reference's download)."""
        if self.z0 != 0 or self.z1 != self.dims[2]:
            raise _lib.DfuError(1, "marching cubes needs the whole volume, this object holds a z-slab")
        cap = int(capacity) if capacity is not None else 6 * 1000 * 1000  # DEFAULT_TRIANGLES_BUFFER_SIZE (marching_cubes.hpp:22)
        verts = torch.empty((cap, 4), dtype=torch.float32, device=self.device)
        ids = torch.empty(cap, dtype=torch.int32, device=self.device) if with_cube_ids else None
        n = torch.zeros(1, dtype=torch.int32, device=self.device)
        check(lib.dfu_marching_cubes(self._base_ptr(), iarr(self.dims), farr(self.size), dptr(verts), dptr(ids), cap, dptr(n),
                                     stream_ptr(device=self.device)))
        total = int(n.item())
        m = min(total, cap)
        if with_cube_ids:
            return verts[:m], ids[:m], total
        return verts[:m]

    # TsdfVolume::raycast (tsdf_volume.cpp:95-129): points or depth + normals of the fused model seen from camera_pose
    raycast_step_factor = 0.75    # tsdf_volume.cpp:26
    gradient_delta_factor = 0.75  # tsdf_volume.cpp:25 (KinFuParams sets 0.5, kinfu.cpp:38)

    def setRaycastStepFactor(self, f):
        self.raycast_step_factor = float(f)

    def setGradientDeltaFactor(self, f):
        self.gradient_delta_factor = float(f)

    def raycast(self, camera_pose, intr, rows, cols, want="points"):
        """returns (points [rows, cols, 4] float32 | depth [rows, cols] int16 (uint16 bits), normals [rows, cols, 4])"""
        if self.z0 != 0 or self.z1 != self.dims[2]:
            raise _lib.DfuError(1, "raycast needs the whole volume, this object holds a z-slab")
        cam = torch.as_tensor(camera_pose, dtype=torch.float64).reshape(4, 4)
        cam2vol = torch.linalg.inv(self.pose) @ cam  # pose_.inv() * camera_pose
        rinv = torch.linalg.inv(cam2vol[:3, :3])
        c2v = [float(x) for x in cam2vol[:3, :3].reshape(-1)] + [float(x) for x in cam2vol[:3, 3]]
        ri = [float(x) for x in rinv.reshape(-1)]
        normals = torch.empty((rows, cols, 4), dtype=torch.float32, device=self.device)
        points = depth = None
        if want == "points":
            points = torch.empty((rows, cols, 4), dtype=torch.float32, device=self.device)
        else:
            depth = torch.empty((rows, cols), dtype=torch.int16, device=self.device)
        check(lib.dfu_tsdf_raycast(self._base_ptr(), iarr(self.dims), farr(self.getVoxelSize()), self.trunc_dist, farr(c2v), farr(ri),
                                   farr(intr), rows, cols, self.raycast_step_factor, self.gradient_delta_factor,
                                   dptr(points), cols * 16, dptr(depth), cols * 2, dptr(normals), cols * 16, stream_ptr(device=self.device)))
        return (points if want == "points" else depth), normals
