"""Data-parallel sharding of the hot path over the GPUs of one node (one process per GPU, torch.distributed).

* the TSDF volume is cut into z-slabs of whole 8-plane bricks: rank r owns planes [z0, z1);
* the surface points are cut into contiguous partitions;
* the solver's per-node normal-equation buffers are summed over ranks (NCCL all-reduce over NVLink).
No collective touches the volume."""
import torch
import torch.distributed as dist

BRICK = 8


def slab_range(rank, world, dz):
    """planes [z0, z1) of rank `rank`: contiguous, whole bricks, covering [0, dz) exactly once"""
    nb = (dz + BRICK - 1) // BRICK
    z0 = min(dz, (nb * rank // world) * BRICK)
    z1 = min(dz, (nb * (rank + 1) // world) * BRICK)
    return z0, z1


def point_range(rank, world, n_points):
    """contiguous partition [p0, p1) of the surface points"""
    return n_points * rank // world, n_points * (rank + 1) // world


def make_allreduce(group=None):
    """all-reduce hook for CombinedSolver.setAllReduce / DynFusion.allreduce"""
    def _allreduce(t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return _allreduce


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class Communicator:
    """NCCL communicator owned by libdynfu_b200.so (created over torch.distributed's rendezvous), so that the
    solver issues its all-reduces from C++ on its own stream -- no Python inside the PCG loop."""

    def __init__(self, device=None):
        import ctypes as C

        from ._lib import check, lib

        rank, world = dist.get_rank(), dist.get_world_size()
        buf = C.create_string_buffer(128)
        if rank == 0:
            check(lib.dfu_comm_unique_id(buf))
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0)
        h = C.c_void_p()
        if device is not None:
            torch.cuda.set_device(device)
        check(lib.dfu_comm_create(C.byref(h), box[0], rank, world))
        self._h = h
        self.rank, self.world = rank, world

    @property
    def handle(self):
        return self._h

    def close(self):
        from ._lib import lib

        if getattr(self, "_h", None):
            lib.dfu_comm_destroy(self._h)
            self._h = None
