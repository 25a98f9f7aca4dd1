#!/usr/bin/env python
"""bench.py -- frames/s of DynamicFusion's per-frame hot path on B200 (BASELINE.json's metric).

One step = one frame of the full per-frame loop of BASELINE configs[2] at the size of configs[1]:
    compute_dists -> warpToLive of the canonical surface points (8-NN + DQB) -> data graph
    -> 5 Gauss-Newton iterations x 10 PCG iterations (Tukey re-weighting each GN iteration)
    -> write-back onto the nodes -> WARPED TSDF integration of the live depth into the 512^3 canonical volume
on a synthetic bending cylinder (4096 nodes, ~76k surface points, 640x480 depth).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
N > 1 runs under torchrun (one rank per GPU, NCCL): the volume is sharded into z-slabs, the surface points are
partitioned, the per-node normal-equation buffers are all-reduced ("strong" scaling: same frame, more GPUs).

Prints ONE JSON line (rank 0).  `value` = frames/s with the frame's inputs already resident in HBM;
`e2e` = the same through the public API with HOST inputs (pinned depth + live points up, energy + node
transforms down, every step).  `roofline` is for the dominant kernel (integrate_kernel) from CUDA events
measured inside this run; `cpu_baseline` is the CPU oracle (restated reference path, kNN through the
reference's own nanoflann when oracle/_ref is present) timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tools import synth  # noqa: E402

DIM = 512
N_THETA, N_Y = 64, 64
EPSILON = 0.0125
KAPPA = 0.005  # bend x += KAPPA*a*(y-cy)^2: the largest displacement (a = 3, |y| = 0.8 m) is 9.6 mm.  The reference evaluates
               # node weights at the WARPED positions against static node positions (dyn_fusion.cpp:196-206), which only
               # tracks deformations well below dg_w (37.5 mm): at 0.01 and above the frame-to-frame loop slowly diverges
               # (measured over 260 frames), at 0.005 it settles into a periodic steady state
GN_ITERS, PCG_ITERS = 5, 10
LAMBDA = 200.0
RING = 4
AMPL = [1.0, 2.0, 3.0, 2.0]  # bend amplitude of frame i is KAPPA*AMPL[i % 4]: consecutive frames differ by KAPPA
ALGO_BYTES_PER_VOXEL = 8     # 4 B read + 4 B write of the ushort2 voxel, full sweep


def make_scene():
    depth0 = synth.cylinder_depth()
    canon = synth.backproject(depth0, synth.INTR)
    pos, dq, dg_w = synth.cylinder_nodes(N_THETA, N_Y, EPSILON)
    depths = [synth.cylinder_depth(kappa=KAPPA * a) for a in AMPL]
    lives = [synth.bend(canon, KAPPA * a) for a in AMPL]
    return dict(depth0=depth0, canon=canon, pos=pos, dq=dq, dg_w=dg_w, depths=depths, lives=lives)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampled every 50 ms in the background; samples are kept with their wall-clock time so that only the
    ones taken inside the timed region are reported."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self, t0=None, t1=None):
        import datetime

        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        rows = []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), f[5:9]))
            except ValueError:
                continue
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        inside = [r for r in rows if t0 is not None and t0 - 0.05 <= r[0] <= t1 + 0.05]
        use = inside if inside else rows
        reasons = set()
        for r in use:
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median([r[1] for r in use])), "sm_max_mhz": max(r[2] for r in use), "reasons": sorted(reasons),
                "samples": len(use), "window": "timed region" if inside else "whole run (timed region shorter than the sampling period)"}


# --------------------------------------------------------------------------------------------------------
# CPU arm: the restated reference path (oracle), all host threads, bounded sample of the same frame
def cpu_frame(scene, budget_planes=64):
    from oracle import pyoracle

    try:
        o = pyoracle.Oracle("nanoflann")
        knn = "reference nanoflann KD-tree (oracle/_ref)"
    except FileNotFoundError:
        o = pyoracle.Oracle("brute")
        knn = "brute-force kNN (oracle/_ref absent)"
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers)
    o.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    pos, dq, dg_w, canon = scene["pos"], scene["dq"], scene["dg_w"], scene["canon"]
    live, depth = scene["lives"][0], scene["depths"][0]
    t0 = time.perf_counter()
    dists = o.compute_dists(depth, synth.INTR)
    warped = o.warp(pos, dq, dg_w, canon)
    prm = pyoracle.default_params(num_iter=GN_ITERS, nonlinear_iter=1, linear_iter=PCG_ITERS, lambda_=LAMBDA,
                                  pcg_tol=0.0, early_out=0)
    t_sol, dq_new, _ = o.solve(pos, dq, dg_w, warped, live, prm)
    t_points = time.perf_counter() - t0
    # warped integration on a bounded sample: every (DIM/budget)-th plane, extrapolated to DIM planes
    vs = synth.voxel_size(DIM)
    tr = o.trunc_dist(synth.TRUNC, vs)
    stride = DIM // budget_planes
    planes = list(range(stride // 2, DIM, stride))
    vol = np.zeros((1, DIM, DIM), np.uint32)  # one plane of scratch, re-used (the C side offsets by z)
    t1 = time.perf_counter()
    for z in planes:
        # shift the base pointer so that plane z lands in the scratch plane
        base = vol.ctypes.data - z * DIM * DIM * 4
        _integrate_plane(o, base, vs, tr, dists, (pos, dq_new, dg_w), z)
    t_planes = time.perf_counter() - t1
    t_int = t_planes * DIM / len(planes)
    total = t_points + t_int
    return dict(frames_per_s=1.0 / total, cores=o.num_threads(), t_points=t_points, t_integrate_est=t_int,
                sample="1 frame: compute_dists + warp of %d points + %dx%d GN/PCG in full, warped integrate timed on %d of %d "
                       "z-planes (every %dth) and extrapolated; kNN = %s" %
                       (len(canon), GN_ITERS, PCG_ITERS, len(planes), DIM, stride, knn))


def _integrate_plane(o, base_ptr, vs, tr, dists, nodes, z):
    import ctypes as C
    from oracle.pyoracle import _f, _f32

    dims = np.array([DIM, DIM, DIM], np.int32)
    pos, dq, dg_w = _f32(nodes[0], (-1, 3)), _f32(nodes[1], (-1, 8)), _f32(nodes[2])
    v2c, intr = _f32(synth.VOL2CAM), _f32(synth.INTR)
    d = np.ascontiguousarray(dists, np.uint16)
    o.lib.orc_tsdf_integrate(C.cast(base_ptr, C.POINTER(C.c_uint32)), dims.ctypes.data_as(C.POINTER(C.c_int32)), _f(vs),
                             tr, synth.MAX_WEIGHT, _f(v2c), _f(intr), d.ctypes.data_as(C.POINTER(C.c_uint16)),
                             d.shape[1] * 2, d.shape[0], d.shape[1], _f(pos), _f(dq), _f(dg_w), pos.shape[0], 0, z, z + 1,
                             None)


def north_star_variant(scene, devs, steps=40, warmup=4):
    """The same frame with the north-star's own data term instead of the reference's energy: point-to-plane residuals, one
    SE(3) increment per node (one cooperative launch for the 5 GN x 10 PCG solve), true dual-quaternion blending in the warp
    and the integrator.  Device-timed like the headline; reported next to it, not instead of it (the reference has no such
    term: parity is against the double-precision oracle, tests/test_gpu_parity.py)."""
    import torch

    import dynfu_b200 as dfu

    def dev(a, dt=torch.float32):
        return torch.as_tensor(np.ascontiguousarray(a)).to(devs, dtype=dt)

    prm = dfu.DynFuParams(kinfuParams=dfu.KinFuParams(volume_dims=(DIM, DIM, DIM)), epsilon=EPSILON, lambda_=LAMBDA,
                          blend_mode=dfu.BLEND_DQB_SUM,
                          solver=dfu.CombinedSolverParameters(numIter=GN_ITERS, nonLinearIter=1, linearIter=PCG_ITERS,
                                                              earlyOut=False, pcgTolerance=0.0))
    df = dfu.DynFusion(prm, device=devs)
    df.init(dev(scene["canon"]), None, nodes=(dev(scene["pos"]), dev(scene["dq"]), dev(scene["dg_w"])))
    df(torch.from_numpy(scene["depth0"].view(np.int16)).pin_memory())
    # normals of the cylinder (axis parallel to y through the volume-frame centre of synth.cylinder_nodes): radial
    c_vol = np.asarray((0.0, 0.0, 2.0), np.float64) - synth.VOLUME_T
    nn = scene["canon"].astype(np.float64) - c_vol
    nn[:, 1] = 0.0
    nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    live_n = dev(nn.astype(np.float32))
    df.solver.setEnergy(df.solver.ENERGY_P2PLANE_SE3)
    depth_dev = [torch.from_numpy(d.view(np.int16)).to(devs) for d in scene["depths"]]
    live_dev = [dev(l) for l in scene["lives"]]
    kp = prm.kinfuParams
    ev = []

    def frame(i, timed):
        dfu.compute_dists(depth_dev[i % RING], kp.intr, out=df._dists)
        df.canonicalWarpedToLive, _ = df.warpCanonical()
        df.solver.initializeProblemInstance(df.canonicalWarpedToLive, live_dev[i % RING], liveNormals=live_n)
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        df.solver.solveAll()
        if timed:
            b.record()
            ev.append((a, b))
        df.volume.integrate(df._dists, df.camera_pose, kp.intr, df.warpfield, prm.blend_mode)

    for i in range(warmup):
        frame(i, False)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(steps):
        frame(warmup + i, True)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    st = df.solver.getStats()
    return {"workload": "the headline frame with the point-to-plane SE(3) data term and dual-quaternion blending "
                        "(DFU_ENERGY_P2PLANE_SE3, BLEND_DQB_SUM); inputs resident in HBM",
            "value": 1e3 / ms, "unit": "frames/s", "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "solve_ms": float(np.mean([a.elapsed_time(b) for a, b in ev])),
            "solver": {"final_energy": st["final_energy"], "initial_energy": st["initial_energy"],
                       "pcg_iterations": st["pcg_iterations"], "gn_steps": st["gn_steps"]}}



def cpu_variant_solve(scene):
    """CPU baseline of the variant's solve: the double-precision oracle restatement of the point-to-plane SE(3) Gauss-Newton
    / PCG (all host threads) on the first frame's problem -- the stand-in for the Ceres path the north star names, which
    the reference never links."""
    from oracle import pyoracle

    try:
        o = pyoracle.Oracle("nanoflann")
    except FileNotFoundError:
        o = pyoracle.Oracle("brute")
    o.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    c_vol = np.asarray((0.0, 0.0, 2.0), np.float64) - synth.VOLUME_T
    nn = scene["canon"].astype(np.float64) - c_vol
    nn[:, 1] = 0.0
    nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    prm = pyoracle.default_params(num_iter=GN_ITERS, nonlinear_iter=1, linear_iter=PCG_ITERS, lambda_=LAMBDA, pcg_tol=0.0, early_out=0)
    t0 = time.perf_counter()
    _, _, st = o.solve_p2plane(scene["pos"], scene["dq"], scene["dg_w"], scene["canon"], scene["lives"][0], nn.astype(np.float32), prm)
    dt = time.perf_counter() - t0
    return {"solve_ms": dt * 1e3, "cores": o.num_threads(), "kind": "port",
            "sample": "the %dx%d point-to-plane solve of the first frame in full (graphs + weights + GN/PCG, double precision)" %
                      (GN_ITERS, PCG_ITERS), "final_energy": float(st[1])}


def config_dict(n_gpus, P):
    return {"workload": "C3 full per-frame loop at C2 size: compute_dists + warpToLive(8-NN+DQB) + %d GN x %d PCG + warped "
                        "TSDF integrate, %d^3 volume, %d nodes, %d surface points, 640x480 depth, bending cylinder" %
                        (GN_ITERS, PCG_ITERS, DIM, N_THETA * N_Y, P),
            "volume": "%d^3 ushort2 (512 MiB, larger than the 126 MB L2: no L2 flush needed between steps)" % DIM,
            "blend": "REF_COMPOSE (reference calcDQB)", "lambda": LAMBDA, "parallelism":
            "z-slabs x point partitions over %d GPU(s)" % n_gpus}


def run_reference(args):
    """--impl reference: the restated reference path (oracle; kNN through the reference's own nanoflann) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    scene = make_scene()
    for _ in range(min(args.warmup, 1)):
        cpu_frame(scene, budget_planes=4)
    vals = []
    t_start = time.time()
    for _ in range(max(1, args.steps)):
        vals.append(cpu_frame(scene, budget_planes=64))
        if time.time() - t_start > 150.0:  # keep the whole run within a few minutes whatever the host
            break
    best = max(vals, key=lambda r: r["frames_per_s"])
    fps = float(len(vals) / sum(1.0 / r["frames_per_s"] for r in vals))
    line = {"impl": "reference", "metric": "frames/sec (512^3 TSDF, 4096 nodes)", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": len(vals), "steps_requested": args.steps, "warmup": min(args.warmup, 1),
            "ms_per_step": 1000.0 / fps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.gpus, len(scene["canon"])),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": best["cores"], "kind": "port",
                             "sample": best["sample"]},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------------
def parity_check(scene, devs):
    """first frame of the bench scene solved on the GPU and by the CPU oracle: energies and node translations"""
    import torch

    import dynfu_b200 as dfu
    from oracle import pyoracle

    def dev(a):
        return torch.as_tensor(np.ascontiguousarray(a)).to(devs, dtype=torch.float32)

    try:
        o = pyoracle.Oracle("nanoflann")
    except FileNotFoundError:
        o = pyoracle.Oracle("brute")
    o.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    prm_o = pyoracle.default_params(num_iter=GN_ITERS, nonlinear_iter=1, linear_iter=PCG_ITERS, lambda_=LAMBDA, pcg_tol=0.0, early_out=0)
    t_o, _, st_o = o.solve(scene["pos"], scene["dq"], scene["dg_w"], scene["canon"], scene["lives"][0], prm_o)
    wf = dfu.Warpfield(devs)
    wf.init(EPSILON, dev(scene["pos"]), dev(scene["dq"]), dev(scene["dg_w"]))
    prm = dfu.CombinedSolverParameters(numIter=GN_ITERS, nonLinearIter=1, linearIter=PCG_ITERS, earlyOut=False, pcgTolerance=0.0)
    s = dfu.CombinedSolver(wf, prm, 4.652, 1e-2, LAMBDA, 1e-4)
    s.initializeProblemInstance(dev(scene["canon"]), dev(scene["lives"][0]))
    s.solveAll()
    st = s.getStats()
    t_g = s.getTranslations().cpu().numpy().astype(np.float64)
    rel_e = abs(st["final_energy"] - st_o[1]) / st_o[1]
    rel_t = float(np.max(np.abs(t_g - t_o)) / np.abs(t_o).max())
    return {"what": "first frame of this scene (identity field -> live frame 0, %d x %d): GPU solve vs the CPU oracle" % (GN_ITERS, PCG_ITERS),
            "final_energy_gpu": st["final_energy"], "final_energy_cpu": float(st_o[1]), "rel_err": rel_e,
            "max_translation_rel_err": rel_t, "tolerance": 1e-4, "ok": bool(rel_e <= 1e-4 and rel_t <= 1e-4)}


def knn_rooflines(scene, devs, fp32_peak):
    """kNN of the surface points among the nodes, timed alone: the grid-bucketed default and the TMA-staged brute force that
    SURVEY 8(d) puts on the FP32 roofline (8 flop per (query, node) pair)"""
    import torch

    import dynfu_b200 as dfu

    def dev(a):
        return torch.as_tensor(np.ascontiguousarray(a)).to(devs, dtype=torch.float32)

    q = dev(scene["canon"])
    out = []
    for kind in ("grid", "brute"):
        if kind == "brute":
            os.environ["DFU_POINT_KNN"] = "brute"
        try:
            wf = dfu.Warpfield(devs)
            wf.init(EPSILON, dev(scene["pos"]), dev(scene["dq"]), dev(scene["dg_w"]))
        finally:
            os.environ.pop("DFU_POINT_KNN", None)
        for _ in range(3):
            wf.findNeighborsIndex(8, q)
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            wf.findNeighborsIndex(8, q)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        flops = 8.0 * len(scene["pos"]) * len(scene["canon"])
        ach = flops / (ms * 1e-3) / 1e12
        out.append({"kernel": "knn8 of %d surface points among %d nodes, %s" %
                              (len(scene["canon"]), len(scene["pos"]),
                               "grid-bucketed exact search (points_grid_kernel, default)" if kind == "grid" else
                               "brute force, node tiles staged by TMA bulk copies (points_kernel)"),
                    "bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach / fp32_peak if fp32_peak else None,
                    "traffic": None, "peak_source": "measured in this run (dfu_microbench_fp32: FFMA chains)", "kernel_ms": ms,
                    "algorithmic_flops_per_launch": flops,
                    "note": "algorithmic flops = 8 per (query, node) pair for ALL pairs (SURVEY 8d)" +
                            ("; the grid search examines ~40 candidates per query instead of %d, so `achieved` is not a rate of "
                             "executed flops" % len(scene["pos"]) if kind == "grid" else "")})
    return out


def dense_integration(scene, devs, hbm_peak, peak_src, steps=40, warmup=6):
    """BASELINE configs[1] with a depth image that has a value at EVERY pixel (the cylinder in front of a wall at 3 m): warped
    integration into the 512^3 volume through the solved warp field.  The whole frustum in front of the surfaces is updated,
    so this is the streaming regime of the integrator; algorithmic bytes = 8 B x voxels actually updated."""
    import ctypes as C

    import torch

    import dynfu_b200 as dfu
    from dynfu_b200._lib import lib

    def dev(a, dt=torch.float32):
        return torch.as_tensor(np.ascontiguousarray(a)).to(devs, dtype=dt)

    prm = dfu.DynFuParams(kinfuParams=dfu.KinFuParams(volume_dims=(DIM, DIM, DIM)), epsilon=EPSILON, lambda_=LAMBDA,
                          solver=dfu.CombinedSolverParameters(numIter=GN_ITERS, nonLinearIter=1, linearIter=PCG_ITERS,
                                                              earlyOut=False, pcgTolerance=0.0))
    df = dfu.DynFusion(prm, device=devs)
    df.init(dev(scene["canon"]), None, nodes=(dev(scene["pos"]), dev(scene["dq"]), dev(scene["dg_w"])))
    df(torch.from_numpy(synth.with_wall(scene["depth0"]).view(np.int16)).pin_memory())
    df.warpCanonicalToLiveOpt(dev(scene["lives"][0]))  # a solved (translation-only) field, like the headline frame
    kp = prm.kinfuParams
    dd = [dfu.compute_dists(dev(synth.with_wall(d).view(np.int16), torch.int16), kp.intr) for d in scene["depths"]]
    for i in range(warmup):
        df.volume.integrate(dd[i % RING], df.camera_pose, kp.intr, df.warpfield, prm.blend_mode)
    torch.cuda.synchronize()
    ev = []
    for i in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        df.volume.integrate(dd[i % RING], df.camera_pose, kp.intr, df.warpfield, prm.blend_mode)
        b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
    st = (C.c_ulonglong * 4)()
    lib.dfu_tsdf_integrate_stats(st, None)
    touched, quads = int(st[0]), int(st[1])
    ach = touched * ALGO_BYTES_PER_VOXEL / (ms * 1e-3) / 1e9
    return {"kernel": "integrate, dense depth (cylinder in front of a wall at 3 m: every pixel has a depth), 512^3, 4096 nodes, warped",
            "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
            "peak_source": peak_src, "kernel_ms": ms, "voxels_updated": touched,
            "algorithmic_bytes_per_launch": touched * ALGO_BYTES_PER_VOXEL,
            "quad_bytes_per_launch": quads * 32, "quad_GBps": quads * 32 / (ms * 1e-3) / 1e9,
            "full_sweep_GBps": DIM ** 3 * ALGO_BYTES_PER_VOXEL / (ms * 1e-3) / 1e9,
            "bricks_warped": int(st[3]),
            "note": "algorithmic bytes = 8 B per UPDATED voxel (4 B read + 4 B write; the reference also only touches updated "
                    "voxels, tsdf_volume.cu:79-90); quad bytes = the 16-byte quads actually read and written (a quad with one updated "
                    "voxel moves all four); full_sweep = 8 B x every voxel of the volume / time, for comparison with SURVEY 8(d)"}


def multi_gpu_subrecords(rank, world, devs, comm, reps=5):
    """N > 1: the two configurations of BASELINE.json that DO shard (the 0.5 ms headline frame cannot strong-scale: its
    solve is latency-bound and replicated), each next to the same work on ONE GPU of the same run:
      c5_partitioned  configs[4]: 300 k surface points / 32 768 nodes, 10 GN x 10 PCG -- points partitioned over the ranks,
                      per-node normal-equation blocks all-reduced over NVLink by NCCL (one all-reduce per GN step + one per
                      PCG iteration, issued by the library on the solver's stream);
      c4_slab         configs[3]: warped integration into a 1024^3 volume with 16 384 nodes, one z-slab per rank (no
                      collective on the data path).
    Times are CUDA-event medians, max over ranks; `single_gpu_ms` is rank 0 doing ALL the work alone while the others wait."""
    import torch
    import torch.distributed as dist

    import dynfu_b200 as dfu
    from dynfu_b200 import dist as dfu_dist

    def dev(a, dt=torch.float32):
        return torch.as_tensor(np.ascontiguousarray(a)).to(devs, dtype=dt)

    def median_ms(fn, reset=None, n=reps, warm=2):
        ts = []
        for i in range(warm + n):
            if reset is not None:
                reset()
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            if i >= warm:
                ts.append(a.elapsed_time(b))
        return dfu_dist.max_over_ranks(float(np.median(ts)), devs)

    out = {}
    # ---- C5: partitioned data-term solve ------------------------------------------------------------------------
    try:
        eps = 0.004
        pos, dq, dg_w = synth.cylinder_nodes(256, 128, eps)
        rng = np.random.default_rng(synth.SEED + 8)
        th = np.pi + rng.uniform(0.0, np.pi, 300000)
        c_vol = np.array([0.0, 0.0, 2.0]) - synth.VOLUME_T
        canon = np.stack([c_vol[0] + 0.3 * np.cos(th), c_vol[1] + rng.uniform(-0.8, 0.8, 300000), c_vol[2] + 0.3 * np.sin(th)],
                         -1).astype(np.float32)
        live = synth.bend(canon, 0.005 * eps / 0.0125)
        prm = dfu.CombinedSolverParameters(numIter=10, nonLinearIter=1, linearIter=10, earlyOut=False, pcgTolerance=0.0)
        ident = dev(dq)

        def make(p0, p1, with_comm):
            wf = dfu.Warpfield(devs)
            wf.init(eps, dev(pos), ident, dev(dg_w))
            sv = dfu.CombinedSolver(wf, prm, 4.652, 1e-2, LAMBDA, 1e-4)
            if with_comm:
                sv.setCommunicator(comm)
            sv.initializeProblemInstance(dev(canon[p0:p1]), dev(live[p0:p1]))
            return wf, sv

        p0, p1 = dfu_dist.point_range(rank, world, len(canon))
        wf, sv = make(p0, p1, True)
        ms_n = median_ms(sv.solveAll, reset=lambda: wf.setTransformations(ident))
        st_n = sv.getStats()
        del sv, wf
        ms_1, st_1 = None, None
        if rank == 0:
            wf, sv = make(0, len(canon), False)
        # (every rank takes part in the barriers of median_ms; only rank 0 works)
        ms_1 = median_ms((lambda: sv.solveAll()) if rank == 0 else (lambda: None),
                         reset=(lambda: wf.setTransformations(ident)) if rank == 0 else None)
        if rank == 0:
            st_1 = sv.getStats()
            del sv, wf
        t1 = torch.tensor([ms_1 if rank == 0 else 0.0], device=devs)
        dist.broadcast(t1, 0)
        ms_1 = float(t1.item())
        if rank == 0:
            out["c5_partitioned"] = {
                "workload": "configs[4]: 300000 surface points, 32768 nodes, 10 GN x 10 PCG, lambda %g; points partitioned over %d "
                            "ranks" % (LAMBDA, world),
                "collective": "ncclAllReduce(sum, f32) of [J^T r | diag J^T J | E] (4N+4 floats) per GN step and of A p (3N floats) per "
                              "PCG iteration, on the solver's stream (comm.cu)",
                "ms": ms_n, "single_gpu_ms": ms_1, "speedup": ms_1 / ms_n, "efficiency": ms_1 / ms_n / world,
                "solves_per_s": 1e3 / ms_n, "final_energy": st_n["final_energy"], "final_energy_single_gpu": st_1["final_energy"],
                "energy_rel_diff": abs(st_n["final_energy"] - st_1["final_energy"]) / st_1["final_energy"],
                "note": "single_gpu_ms is the one-launch persistent kernel (no exchange); the partitioned path runs one kernel per "
                        "phase between its all-reduces"}
    except Exception as e:
        if rank == 0:
            out["c5_partitioned"] = {"error": repr(e)}
    torch.cuda.empty_cache()
    # ---- C4: z-slab sharded warped integration ------------------------------------------------------------------
    try:
        dim, eps = 1024, 0.00625
        depth0 = synth.cylinder_depth()
        canon = synth.backproject(depth0, synth.INTR)
        pos, dq, dg_w = synth.cylinder_nodes(128, 128, eps)
        kappa = 0.005 * eps / 0.0125
        depth1 = synth.cylinder_depth(kappa=kappa)
        live = synth.bend(canon, kappa)
        prm = dfu.DynFuParams(kinfuParams=dfu.KinFuParams(volume_dims=(dim, dim, dim)), epsilon=eps, lambda_=LAMBDA,
                              solver=dfu.CombinedSolverParameters(numIter=GN_ITERS, nonLinearIter=1, linearIter=PCG_ITERS,
                                                                  earlyOut=False, pcgTolerance=0.0))

        def slab_run(z0, z1, active):
            if not active:
                return median_ms(lambda: None, n=reps, warm=3), 0
            df = dfu.DynFusion(prm, device=devs, z0=z0, z1=z1)
            df.init(dev(canon), None, nodes=(dev(pos), dev(dq), dev(dg_w)))
            df(torch.from_numpy(depth0.view(np.int16)).pin_memory())
            df.warpCanonicalToLiveOpt(dev(live))  # replicated solve: every rank holds the same field
            d1 = dfu.compute_dists(dev(depth1.view(np.int16), torch.int16), prm.kinfuParams.intr)
            ms = median_ms(lambda: df.volume.integrate(d1, df.camera_pose, prm.kinfuParams.intr, df.warpfield, prm.blend_mode),
                           n=reps, warm=3)
            import ctypes as C
            from dynfu_b200._lib import lib
            st = (C.c_ulonglong * 4)()
            lib.dfu_tsdf_integrate_stats(st, None)
            del df
            torch.cuda.empty_cache()
            return ms, int(st[0])

        z0, z1 = dfu_dist.balanced_slab_range(rank, world, dim, pos[:, 2], 3.0 / dim, 0.25, behind=0.06)
        ms_n, vox_n = slab_run(z0, z1, True)
        tv = torch.tensor([float(vox_n)], device=devs, dtype=torch.float64)
        dist.all_reduce(tv, op=dist.ReduceOp.SUM)
        ms_1, vox_1 = slab_run(0, dim, rank == 0)
        t1 = torch.tensor([ms_1 if rank == 0 else 0.0, float(vox_1)], device=devs, dtype=torch.float64)
        dist.broadcast(t1, 0)
        ms_1, vox_1 = float(t1[0].item()), int(t1[1].item())
        if rank == 0:
            out["c4_slab"] = {
                "workload": "configs[3]: warped integration into a 1024^3 volume (4 GiB), 16384 nodes, 640x480 depth; %d z-slabs cut "
                            "for equal work" % world,
                "collective": "none on the data path (node transforms are replicated: every rank runs the same deterministic solve)",
                "ms": ms_n, "single_gpu_ms": ms_1, "speedup": ms_1 / ms_n, "efficiency": ms_1 / ms_n / world,
                "voxels_updated_all_ranks": int(tv.item()), "voxels_updated_single_gpu": vox_1,
                "voxels_per_s": dim ** 3 / (ms_n * 1e-3)}
    except Exception as e:
        if rank == 0:
            out["c4_slab"] = {"error": repr(e)}
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--loops", type=int, default=0, help="repetitions of the timed K-step loop (0: as many as fit in ~1.5 s, 5..50)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variant", action="store_true", help="skip the sub-records (point-to-plane variant, dense integration, kNN)")
    ap.add_argument("--no-overlap", action="store_true", help="integrate on the same stream as the solve (sequential schedule)")
    ap.add_argument("--solver-parallel", default="auto", choices=["auto", "replicated", "partitioned"],
                    help="N > 1: every rank solves all points (no exchange) or the points are partitioned and the "
                         "normal-equation buffers all-reduced over NCCL; auto = partitioned from 100k points per rank")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import ctypes as C

    import torch
    import torch.distributed as dist

    import dynfu_b200 as dfu
    from dynfu_b200._lib import lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU: the product has no CPU fallback"
    torch.cuda.set_device(local)
    devs = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=devs)

    scene = make_scene()
    P_all = len(scene["canon"])
    from dynfu_b200 import dist as dfu_dist
    # slabs of equal WORK: the near bricks cluster around the surface (a 0.6 m thick band of the 3 m volume)
    z0, z1 = dfu_dist.balanced_slab_range(rank, world, DIM, scene["pos"][:, 2], 3.0 / DIM, 0.25, behind=0.06) if world > 1 else (0, DIM)
    # the solve is latency-bound at this size (76k points, 4096 nodes): partitioning the points only adds one
    # all-reduce per PCG iteration, so by default every rank solves the whole (small) problem and only the volume
    # is sharded; the partitioned + all-reduce mode is what larger problems (BASELINE configs[4]) use -- see `c5_partitioned`
    mode = args.solver_parallel
    if mode == "auto":
        mode = "partitioned" if (world > 1 and P_all // world >= 100000) else "replicated"
    if world == 1:
        mode = "single"
    p0, p1 = dfu_dist.point_range(rank, world, P_all) if mode == "partitioned" else (0, P_all)
    overlap = not args.no_overlap

    def dev(a, dt=torch.float32):
        return torch.as_tensor(np.ascontiguousarray(a)).to(devs, dtype=dt)

    pc = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            pc = parity_check(scene, devs)
        except Exception as e:  # the headline line must not depend on the checker
            pc = {"error": repr(e)}

    prm = dfu.DynFuParams(kinfuParams=dfu.KinFuParams(volume_dims=(DIM, DIM, DIM)), epsilon=EPSILON, lambda_=LAMBDA,
                          solver=dfu.CombinedSolverParameters(numIter=GN_ITERS, nonLinearIter=1, linearIter=PCG_ITERS,
                                                              earlyOut=False, pcgTolerance=0.0))
    df = dfu.DynFusion(prm, device=devs, z0=z0, z1=z1)
    df.stream_overlap = overlap
    comm = None
    if world > 1:
        comm = dfu_dist.Communicator(devs)
    if mode == "partitioned":
        df.comm = comm
    df.init(dev(scene["canon"][p0:p1]), None, nodes=(dev(scene["pos"]), dev(scene["dq"]), dev(scene["dg_w"])))

    depth_host = [torch.from_numpy(d.view(np.int16)).pin_memory() for d in scene["depths"]]
    live_host = [torch.from_numpy(np.ascontiguousarray(l[p0:p1])).pin_memory() for l in scene["lives"]]
    depth_dev = [d.to(devs) for d in depth_host]
    live_dev = [l.to(devs) for l in live_host]

    # frame 0 (untimed): the canonical volume, rigid integration of the undeformed surface
    df(torch.from_numpy(scene["depth0"].view(np.int16)).pin_memory())
    torch.cuda.synchronize()

    timers = {}

    def step_device(i, timed=False):
        """inputs already in HBM"""
        df.frameDevice(depth_dev[i % RING], live_dev[i % RING], overlap=overlap, timers=timers if timed else None)

    def step_host(i):
        """public API with HOST buffers (DynFusion.streamFrame): depth + live points up from pinned memory, node transforms +
        solver statistics down to pinned memory, every step; the transfers of step i overlap the kernels of its neighbours
        (two staging slots, separate H2D / D2H copy streams), the host consumes the result of step i-1"""
        return df.streamFrame(depth_host[i % RING], live_host[i % RING])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cpu_submit = {}

    def timed_loop(fn, K, fin=None, **kw):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_cpu = time.perf_counter()
        a.record()
        for i in range(K):
            fn(i, **kw)
        if fin is not None:
            fin()  # (pipelined loops: the last frame's integration / results are inside the timed region)
        b.record()
        cpu_submit[fn.__name__] = (time.perf_counter() - t_cpu) * 1e3 / K  # host time to enqueue one step
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=devs)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(args.warmup):
        step_device(i)
    df.frameSync()
    for i in range(args.warmup):
        step_host(i)
    df.streamFlush()
    barrier()

    # EXACTLY K steps per loop, bracketed by barrier + synchronize, max over ranks; the loop is repeated and `value` is the
    # median loop (a single 20-step loop is a 10 ms sample)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)  # let nvidia-smi start sampling
    launches0 = lib.dfu_launch_count()
    t_wall0 = time.time()
    first_ms = timed_loop(step_device, args.steps, fin=df.frameSync, timed=True)
    launches = lib.dfu_launch_count() - launches0
    loops = args.loops if args.loops > 0 else int(min(50, max(5, 1500.0 / max(first_ms, 1e-3))))
    if world > 1:  # every rank must run the same number of loops
        t = torch.tensor([loops], device=devs)
        dist.broadcast(t, 0)
        loops = int(t.item())
    dev_ms = [first_ms] + [timed_loop(step_device, args.steps, fin=df.frameSync, timed=True) for _ in range(loops - 1)]
    e2e_ms = [timed_loop(step_host, args.steps, fin=df.streamFlush) for _ in range(loops)]
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    stats = df.solver.getStats()
    ms_dev = float(np.median(dev_ms))
    ms_e2e = float(np.median(e2e_ms))

    def spread(v):
        v = np.asarray(v) / args.steps
        return {"loops": int(len(v)), "ms_per_step_median": float(np.median(v)), "min": float(v.min()), "max": float(v.max()),
                "p10": float(np.quantile(v, 0.1)), "p90": float(np.quantile(v, 0.9)), "first_loop": float(v[0])}

    # kernels, timed with CUDA events on the stream they are launched on, inside the timed loops
    int_ms = dfu_dist.max_over_ranks(float(np.median([a.elapsed_time(b) for a, b in timers["integrate"]])), devs)
    sol_ms = dfu_dist.max_over_ranks(float(np.median([a.elapsed_time(b) for a, b in timers["solve"]])), devs)
    ist = (C.c_ulonglong * 4)()
    lib.dfu_tsdf_integrate_stats(ist, None)
    touched = torch.tensor([float(ist[0]), float(ist[1])], device=devs, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(touched, op=dist.ReduceOp.MAX)  # per-rank figures: the slowest rank's launch is what int_ms times
    touched_vox, quads = int(touched[0].item()), int(touched[1].item())
    voxels_rank = DIM * DIM * (z1 - z0)
    hbm_peak, peak_src = peaks()
    achieved = touched_vox * ALGO_BYTES_PER_VOXEL / (int_ms * 1e-3) / 1e9
    # per GN iteration the assembly reads ~100 B/point, per PCG iteration ~100 B/point + N*(6+27)*4 B
    P_rank = p1 - p0
    sol_bytes = GN_ITERS * P_rank * 100 + GN_ITERS * PCG_ITERS * (P_rank * 100 + N_THETA * N_Y * 33 * 4)
    sol_achieved = sol_bytes / (sol_ms * 1e-3) / 1e9

    if rank == 0:
        fps = args.steps / (ms_dev * 1e-3)
        fps_e2e = args.steps / (ms_e2e * 1e-3)
        step_ms = ms_dev / args.steps
        traffic = {}
        tf = os.path.join(ROOT, "profiles", "traffic_r02.json")
        if world == 1 and os.path.exists(tf):  # one ncu --set full capture of THIS command at N = 1 (never re-used for slabs)
            try:
                traffic = json.load(open(tf))
            except Exception:
                traffic = {}
        line = {
            "metric": "frames/sec (512^3 TSDF, 4096 nodes)", "value": fps, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config_dict(world, P_all), solver_parallel=mode,
                           schedule="integrate(i) on a second stream, overlapping the point pipeline of frame i+1" if overlap else
                                    "sequential (one stream)",
                           timing="K steps per loop (barrier + synchronize on both sides, CUDA events, max over ranks); "
                                  "value = K / median over `timing.loops` loops"),
            "timing": spread(dev_ms),
            "e2e": {"value": fps_e2e, "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(depth_host[0].numel() * 2 + live_host[0].numel() * 4),
                    "d2h_bytes_per_step": int(N_THETA * N_Y * 8 * 4 + 32),
                    "pipelined": "2 staging slots, H2D / D2H on copy streams; result of step i is consumed during step i+1",
                    "timing": spread(e2e_ms)},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": cpu_submit.get("step_device"),
            "voxels_per_s": DIM ** 3 * fps,
            "roofline": None,
            "roofline_kernels": [
                {"kernel": "integrate (depth_tiles + tile_classify + integrate_kernel<FILL> + integrate_kernel<CACHED>): "
                           "warped projective TSDF integration", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                 "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic.get("integrate_dram_bytes_per_launch"),
                 "peak_source": peak_src, "kernel_ms": int_ms,
                 "algorithmic_bytes_per_launch": touched_vox * ALGO_BYTES_PER_VOXEL, "voxels_updated": touched_vox,
                 "quad_bytes_per_launch": quads * 32,
                 "full_sweep": {"bytes": voxels_rank * ALGO_BYTES_PER_VOXEL,
                                "GBps": voxels_rank * ALGO_BYTES_PER_VOXEL / (int_ms * 1e-3) / 1e9,
                                "note": "8 B x EVERY voxel of the rank's slab / time (SURVEY 8d's sweep model); not a roofline "
                                        "fraction -- the kernel proves most voxels untouched and never loads them"},
                 "share_of_step": int_ms / step_ms,
                 "note": "algorithmic bytes = 8 B x voxels UPDATED by this launch (counted in the kernel: 4 B read + 4 B write of "
                         "the ushort2; the reference touches the same voxels, tsdf_volume.cu:79-90).  This scene has depth on a "
                         "quarter of the image, so the launch is latency-bound (per-voxel neighbour cache), not streaming: see "
                         "`integrate_dense` for the bandwidth regime"},
                {"kernel": "k_solve_persistent3r: explicit normal matrix + pipelined PCG, 5 GN x 10 PCG in one cooperative launch"
                           if mode != "partitioned" else "solver phase kernels + NCCL all-reduces (5 GN x 10 PCG)", "bound": "hbm", "achieved": sol_achieved,
                 "peak": hbm_peak, "unit": "GB/s", "frac": sol_achieved / hbm_peak, "traffic": traffic.get("solver_dram_bytes_per_launch"),
                 "peak_source": peak_src,
                 "kernel_ms": sol_ms, "algorithmic_bytes_per_launch": sol_bytes, "share_of_step": sol_ms / step_ms,
                 "note": "L2-resident and bound by inter-SM exchange latency (one L2 round trip pair per PCG iteration), not by HBM "
                         "(SURVEY 8d); frac is reported for completeness -- profiles/r02_solver_experiments.md"}],
            "solver": {"final_energy": stats["final_energy"], "initial_energy": stats["initial_energy"],
                       "pcg_iterations": stats["pcg_iterations"], "gn_steps": stats["gn_steps"]},
            "parity_check": pc,
            "clocks": clocks,
        }
        if traffic:
            line["traffic_source"] = traffic.get("source")
        for rk in line["roofline_kernels"]:
            if rk["traffic"]:
                rk["dram_achieved_GBps"] = rk["traffic"] / (rk["kernel_ms"] * 1e-3) / 1e9
        if overlap:
            line["roofline_kernels"][0]["note"] += ("; with the overlapped schedule the kernel shares the SMs with the next frame's point "
                                                    "pipeline, so kernel_ms is its duration on its own stream, not its share of the step")
        # the contract's `roofline` object is the dominant kernel of the step
        line["roofline"] = max(line["roofline_kernels"], key=lambda r: r["kernel_ms"])
    del df
    torch.cuda.empty_cache()
    if world == 1 and not args.no_variant and rank == 0:
        try:
            tf_, mhz = C.c_double(), C.c_double()
            lib.dfu_microbench_fp32(local, C.byref(tf_), C.byref(mhz))
            line["fp32_peak"] = {"tflops": tf_.value, "implied_sm_mhz": mhz.value, "how": "dfu_microbench_fp32: 16 FFMA chains per thread, "
                                 "8 CTAs of 256 threads per SM, best of 4"}
            line["roofline_kernels"] += knn_rooflines(scene, devs, tf_.value)
        except Exception as e:
            line["fp32_peak"] = {"error": repr(e)}
        try:
            line["integrate_dense"] = dense_integration(scene, devs, hbm_peak, peak_src)
            line["roofline_kernels"].append(line["integrate_dense"])
        except Exception as e:
            line["integrate_dense"] = {"error": repr(e)}
        try:
            line["north_star_data_term"] = north_star_variant(scene, devs)
        except Exception as e:  # the headline line must not depend on the extension
            line["north_star_data_term"] = {"error": repr(e)}
        if not args.no_cpu_baseline and "error" not in line["north_star_data_term"]:
            try:
                line["north_star_data_term"]["cpu_baseline"] = cpu_variant_solve(scene)
            except Exception as e:
                line["north_star_data_term"]["cpu_baseline"] = {"error": repr(e)}
    if world > 1 and not args.no_variant:
        sub = multi_gpu_subrecords(rank, world, devs, comm)
        if rank == 0:
            line.update(sub)
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            c = cpu_frame(scene)
            line["cpu_baseline"] = {"value": c["frames_per_s"], "unit": "frames/s", "cores": c["cores"], "kind": "port",
                                    "sample": c["sample"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
