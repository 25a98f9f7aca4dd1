"""world_size-2 gloo tests (CPU) of the data-parallel host logic: slab / point partitioning and the
all-reduce contract of the solver (partition the points, sum the per-node normal-equation blocks)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import synth


def test_slab_and_point_ranges_cover_exactly_once():
    from dynfu_b200.dist import point_range, slab_range

    for dz in (8, 64, 100, 512, 1024):
        for world in (1, 2, 3, 4, 8):
            z = [slab_range(r, world, dz) for r in range(world)]
            assert z[0][0] == 0 and z[-1][1] == dz
            for (a0, a1), (b0, b1) in zip(z, z[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(z0 % 8 == 0 for z0, _ in z)
    for n in (0, 1, 7, 75852):
        for world in (1, 2, 8):
            p = [point_range(r, world, n) for r in range(world)]
            assert p[0][0] == 0 and p[-1][1] == n and all(a[1] == b[0] for a, b in zip(p, p[1:]))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dynfu_b200.dist import make_allreduce, point_range, slab_range
    from oracle import pyoracle

    o = pyoracle.Oracle("brute")
    rng = np.random.default_rng(5)
    pos, _, dg_w, t_true = synth.sphere_nodes(128, 0.05)
    P = 1500
    canon = (pos[rng.integers(0, 128, P)] + rng.normal(0, 0.02, (P, 3))).astype(np.float32)
    live = (canon + 0.01).astype(np.float32)
    p0, p1 = point_range(rank, world, P)

    def blocks(c, l):  # [b | D | E] of energy.t at t = 0 for a set of points (tukey weights at t = 0)
        idx, _ = o.knn(pos, c)
        w = np.array([[o.node_weight(pos[j], dg_w[j], c[v]) for j in idx[v]] for v in range(len(c))], np.float64)
        d = l.astype(np.float64) - c
        th = np.array([o.tukey(4.652, 1e-2, d[v].astype(np.float32)) for v in range(len(c))], np.float64)
        buf = np.zeros(4 * 128 + 4)
        for v in range(len(c)):
            for k in range(8):
                n = idx[v, k]
                buf[3 * n:3 * n + 3] += th[v] * w[v, k] * d[v]
                buf[3 * 128 + n] += th[v] * w[v, k] ** 2
            buf[4 * 128] += th[v] * d[v] @ d[v]
        return buf

    part = torch.from_numpy(blocks(canon[p0:p1], live[p0:p1]))
    make_allreduce()(part)  # the hook bench.py gives the solver
    full = blocks(canon, live)
    ok = bool(np.allclose(part.numpy(), full, rtol=1e-12, atol=1e-15))
    # volume slabs: the oracle integrates [z0,z1) of each rank; together they equal the full volume
    depth = synth.sphere_depth()
    dists = o.compute_dists(depth, synth.INTR)
    dim = 32
    vs = synth.voxel_size(dim)
    z0, z1 = slab_range(rank, world, dim)
    vol = np.zeros((dim,) * 3, np.uint32)
    o.tsdf_integrate(vol, vs, o.trunc_dist(synth.TRUNC, vs), synth.MAX_WEIGHT, synth.VOL2CAM, synth.INTR, dists, z0=z0, z1=z1)
    t = torch.from_numpy(vol.astype(np.int64))
    dist.all_reduce(t)  # slabs are disjoint: the sum is the union
    ref = np.zeros((dim,) * 3, np.uint32)
    o.tsdf_integrate(ref, vs, o.trunc_dist(synth.TRUNC, vs), synth.MAX_WEIGHT, synth.VOL2CAM, synth.INTR, dists)
    ok = ok and bool(np.array_equal(t.numpy().astype(np.uint32), ref))
    out[rank] = ok
    dist.destroy_process_group()


def test_two_rank_gloo_partition_and_allreduce():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_balanced_slabs_tile_the_volume_and_follow_the_nodes():
    import numpy as np
    from dynfu_b200 import dist as dd

    rng = np.random.default_rng(0)
    node_z = rng.normal(1.4, 0.1, 4096)  # a surface band around z = 1.4 m of a 3 m volume
    for world in (1, 2, 3, 4, 8):
        prev = 0
        for r in range(world):
            z0, z1 = dd.balanced_slab_range(r, world, 512, node_z, 3.0 / 512, 0.25)
            assert z0 == prev and z1 > z0 and z0 % 8 == 0 and (z1 % 8 == 0 or z1 == 512)
            prev = z1
        assert prev == 512
    # 2 ranks: the cut goes through the band, not through the middle of the volume (plane 256 = 1.5 m would leave
    # most nodes on one side)
    z0, z1 = dd.balanced_slab_range(0, 2, 512, node_z, 3.0 / 512, 0.25)
    assert abs(z1 * 3.0 / 512 - 1.4) < 0.1
    # degenerate inputs still tile
    assert dd.balanced_slab_range(0, 1, 64, [], 0.05, 0.1) == (0, 64)
    parts = [dd.balanced_slab_range(r, 8, 16, [0.1], 0.05, 0.1) for r in range(8)]
    assert parts[0][0] == 0 and parts[-1][1] == 16 and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
