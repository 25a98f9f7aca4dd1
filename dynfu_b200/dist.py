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


def balanced_slab_range(rank, world, dz, node_z, voxel_z, reach, behind=None, rigid_cost=0.05):
    """z-slabs of whole bricks with (approximately) equal integration work instead of equal thickness.

    The warped integrator spends its time in the bricks within `reach` metres of a deformation node (they run the exact
    per-voxel warp); everything else is the cheap rigid pass or culled.  Work per brick plane ~ number of nodes within
    `reach` in front of the plane and `behind` (default: reach) behind it along +z -- bricks further than the truncation
    band behind the surface are culled, so for a camera looking along +z `behind` is about the truncation distance --
    (+ `rigid_cost` of the mean per plane for the rigid remainder).  Every rank evaluates the same
    deterministic split from the same node positions, so the slabs tile [0, dz) exactly once.  node_z: z coordinates of
    the nodes in volume-local metres (any array-like); voxel_z: voxel edge in metres."""
    import numpy as np

    nb = (dz + BRICK - 1) // BRICK
    z = np.asarray(node_z, dtype=np.float64).reshape(-1)
    centres = (np.arange(nb) * BRICK + 0.5 * (BRICK - 1)) * float(voxel_z)
    hb = 0.5 * BRICK * float(voxel_z)
    behind = float(reach) if behind is None else float(behind)
    zs = np.sort(z)
    # nodes whose z lies in [plane - behind, plane + reach]: the plane is at most `reach` in front of / `behind` behind them
    work = (np.searchsorted(zs, centres + hb + float(reach)) - np.searchsorted(zs, centres - hb - behind)).astype(np.float64)
    work += rigid_cost * max(work.mean(), 1e-12)
    cum = np.concatenate([[0.0], np.cumsum(work)])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cum, cum[-1] * r / world)))
    cuts.append(nb)
    for r in range(1, world + 1):  # monotone, and never an empty slab while there are bricks left
        cuts[r] = max(cuts[r], min(cuts[r - 1] + 1, nb))
    cuts[world] = nb
    return min(dz, cuts[rank] * BRICK), min(dz, cuts[rank + 1] * BRICK)


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
