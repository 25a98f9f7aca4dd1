"""DynFusion -- the per-frame operator of the hot path (kfusion::KinFu::operator() / DynFusion::operator(),
src/kfusion/kinfu.cpp:140-234, src/dynfu/dyn_fusion.cpp:48-145), restricted to the steps SURVEY.md §8 puts in
scope: computeDists -> warpToLive (kNN + DQB) -> CombinedSolver -> warped TSDF integration.

The rows either side of the path (§8f) are wired in as well: live vertices come from the depth image
(cuda::computePointNormals + compaction, in place of the reference's marching-cubes download with its faked
normals, dyn_fusion.cpp:126-134), correspondences from the grid 1-NN (findCorrespondingFrame, :212-242) and the
warp field grows through Warpfield::update (:142).  Marching cubes and ICP stay out (rendering / the reference
skips ICP itself, :100-105); canonical vertices can still be supplied by the caller like the solver tests do."""
from dataclasses import dataclass, field

import numpy as np
import torch

from . import frontend
from ._lib import BLEND_REF_COMPOSE
from .solver import CombinedSolver, CombinedSolverParameters
from .tsdf_volume import TsdfVolume, compute_dists
from .warpfield import Warpfield


@dataclass
class KinFuParams:
    """kfusion::KinFuParams::default_params (src/kfusion/kinfu.cpp:10-44), hot-path fields only."""
    cols: int = 640
    rows: int = 480
    intr: tuple = (525.0, 525.0, 319.5, 239.5)
    volume_dims: tuple = (512, 512, 512)
    volume_size: tuple = (3.0, 3.0, 3.0)
    volume_pose_t: tuple = (-1.5, -1.5, 0.5)  # Affine3f().translate(-size/2, -size/2, 0.5)
    tsdf_trunc_dist: float = 0.04
    tsdf_max_weight: int = 64


@dataclass
class DynFuParams:
    """DynFuParams::defaultParams (src/dynfu/dyn_fusion.cpp:6-31)."""
    kinfuParams: KinFuParams = field(default_factory=lambda: KinFuParams(volume_dims=(128, 128, 128)))
    tukeyOffset: float = 4.652
    lambda_: float = 200.0
    psi_data: float = 0.01
    psi_reg: float = 1e-4
    L: int = 4
    beta: int = 4
    epsilon: float = 0.1
    node_step: int = 128  # every 128th canonical vertex becomes a node (dyn_fusion.cpp:151)
    solver: CombinedSolverParameters = field(default_factory=CombinedSolverParameters)
    blend_mode: int = BLEND_REF_COMPOSE


class DynFusion:
    def __init__(self, params, device=None, z0=0, z1=None):
        self.params = params
        kp = params.kinfuParams
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.volume = TsdfVolume(kp.volume_dims, self.device, z0=z0, z1=z1, size=kp.volume_size)  # kinfu.cpp:49-54
        self.volume.setTruncDist(kp.tsdf_trunc_dist)
        self.volume.setMaxWeight(kp.tsdf_max_weight)
        pose = torch.eye(4, dtype=torch.float64)
        pose[:3, 3] = torch.tensor(kp.volume_pose_t, dtype=torch.float64)
        self.volume.setPose(pose)
        self.camera_pose = torch.eye(4, dtype=torch.float64)  # poses_.back(): ICP is skipped (dyn_fusion.cpp:100-105)
        self.warpfield = None
        self.solver = None
        self.frame_counter = 0
        self.canonicalVertices = None
        self.canonicalNormals = None
        self.canonicalWarpedToLive = None
        self._depth_dev = torch.empty((kp.rows, kp.cols), dtype=torch.int16, device=self.device)
        self._dists = torch.empty_like(self._depth_dev)
        self._points = torch.empty((kp.rows, kp.cols, 4), dtype=torch.float32, device=self.device)
        self._normals = torch.empty_like(self._points)
        self._index = None     # frontend.PointIndex, built per frame like the reference's second KD-tree
        self.liveVertices = None
        self.liveNormals = None
        self.stream_overlap = True  # streamFrame: integrate(i) on a second stream, overlapping the point pipeline of frame i+1
        self.allreduce = None  # python all-reduce hook (tests)
        self.comm = None       # dynfu_b200.dist.Communicator: NCCL issued from C++

    # The canonical frame's nearest nodes and weights are cached on the device, keyed on (pointer, count, version): assigning a
    # new tensor always bumps the version, so a tensor the allocator happens to place at the old address cannot alias the
    # cache.  (Writing INTO the tensor in place is the one thing the key cannot see: call touchCanonical() after doing so.)
    @property
    def canonicalVertices(self):
        return self._canonicalVertices

    @canonicalVertices.setter
    def canonicalVertices(self, value):
        self._canonicalVertices = value
        self._canon_version = getattr(self, "_canon_version", 0) + 1

    def touchCanonical(self):
        self._canon_version += 1

    # DynFusion::init (src/dynfu/dyn_fusion.cpp:147-168); explicit nodes may be given instead of the 128-stride pick
    def init(self, canonicalVertices, canonicalNormals=None, nodes=None):
        cv = torch.as_tensor(canonicalVertices, dtype=torch.float32).to(self.device).reshape(-1, 3).contiguous()
        self.canonicalVertices = cv  # (the setter bumps the version the cached warp below is keyed on)
        self.canonicalNormals = None if canonicalNormals is None else torch.as_tensor(
            canonicalNormals, dtype=torch.float32).to(self.device).reshape(-1, 3).contiguous()
        if nodes is None:
            pos = cv[::self.params.node_step].contiguous()
            n = pos.shape[0]
            dq = torch.zeros((n, 8), dtype=torch.float32, device=self.device)
            dq[:, 0] = 1.0
            w = torch.full((n,), float(np.float32(3) * np.float32(self.params.epsilon)), dtype=torch.float32,
                           device=self.device)  # :156, float arithmetic
        else:
            pos, dq, w = nodes
        self.warpfield = Warpfield(self.device)
        self.warpfield.init(self.params.epsilon, pos, dq, w)
        self.solver = CombinedSolver(self.warpfield, self.params.solver, self.params.tukeyOffset, self.params.psi_data,
                                     self.params.lambda_, self.params.psi_reg)  # dyn_fusion.cpp:193
        if self.comm is not None:
            self.solver.setCommunicator(self.comm)
        elif self.allreduce is not None:
            self.solver.setAllReduce(self.allreduce)

    def uploadDepth(self, depth_host):
        """demo.cpp:90 depth_device_.upload(...): pinned uint16 [rows, cols] -> device."""
        self._depth_dev.copy_(depth_host.view(torch.int16) if depth_host.dtype == torch.uint16 else depth_host,
                              non_blocking=True)
        return self._depth_dev

    # DynFusion::warpCanonicalToLiveOpt (src/dynfu/dyn_fusion.cpp:182-210)
    def warpCanonicalToLiveOpt(self, liveVertices, paired=True):
        """paired=True: liveVertices[i] belongs to canonicalVertices[i] (the solver tests' convention).
        paired=False: the reference's flow -- warp the canonical frame, then pick for every live vertex the nearest
        warped canonical vertex (findCorrespondingFrame, :196-197) and solve on those pairs."""
        self.canonicalWarpedToLive, warpedNormals = self.warpCanonical()
        if paired:
            self.solver.initializeProblemInstance(self.canonicalWarpedToLive, liveVertices)
        else:
            corr_v, _ = self.findCorrespondingFrame(self.canonicalWarpedToLive, warpedNormals, liveVertices)
            self.solver.initializeProblemInstance(corr_v, liveVertices)
        self.solver.solveAll()

    # warpfield->warpToLive(canonicalFrame) (dyn_fusion.cpp:196): the canonical frame is fixed between init() calls, so its
    # nearest nodes and weights are cached on the device (bit-identical to the uncached warp)
    def warpCanonical(self):
        return self.warpfield.warpToLiveCached(self.canonicalVertices, self.canonicalNormals, self._canon_version,
                                               self.params.blend_mode)

    # DynFusion::findCorrespondingFrame (src/dynfu/dyn_fusion.cpp:212-242)
    def findCorrespondingFrame(self, canonicalVertices, canonicalNormals, liveVertices):
        if self._index is None:
            self._index = frontend.PointIndex(self.device)
        return self._index.find_corresponding(canonicalVertices, canonicalNormals, liveVertices)

    # live surface points of a depth frame in the frame of the volume (the nodes' frame): cuda::computePointNormals
    # (kinfu.cpp:170-173) + compaction; replaces the marching-cubes download of dyn_fusion.cpp:120-134
    def liveFrameFromDepth(self, depth_dev):
        kp = self.params.kinfuParams
        frontend.compute_points_normals(depth_dev, kp.intr, self._points, self._normals)
        cam2vol = torch.linalg.inv(self.volume.pose) @ self.camera_pose
        self.liveVertices, self.liveNormals = frontend.compact_points(self._points, self._normals, cam2vol)
        return self.liveVertices, self.liveNormals

    # one fully automatic frame: DynFusion::operator() with the front-end rows on the device
    def processFrame(self, depth_host):
        kp = self.params.kinfuParams
        depth = self.uploadDepth(depth_host)
        compute_dists(depth, kp.intr, out=self._dists)  # :55
        live_v, live_n = self.liveFrameFromDepth(depth)
        if self.frame_counter == 0 or self.warpfield is None:
            self.volume.integrate(self._dists, self.camera_pose, kp.intr)  # :70
            self.init(live_v, live_n)  # :88 (vertices of the first frame are the canonical frame)
            self.frame_counter += 1
            return False
        self.warpCanonicalToLiveOpt(live_v, paired=False)  # :140
        self.warpfield.update(self.canonicalWarpedToLive, self.params.blend_mode)  # :142
        self.volume.integrate(self._dists, self.camera_pose, kp.intr, self.warpfield, self.params.blend_mode)
        self.frame_counter += 1
        return True

    # DynFusion::operator() EXACTLY as the reference runs it (src/dynfu/dyn_fusion.cpp:48-145), for parity runs: the surface
    # points are the marching-cubes vertices of the volume (:74-88, :120-134; "normals" = the positions, FIXME :80), frame 0
    # seeds a node at every 128th vertex (:147-168), every later frame CLEARS the volume and integrates the live depth rigidly
    # (:113-116 -- the reference never fuses non-rigidly), pairs live and warped canonical vertices by 1-NN (:212-242),
    # solves, and grows the warp field (:142).  processFrame() above is the pipeline the north star asks for instead
    # (warped fusion into the canonical volume, live points straight from the depth image).
    def processFrameReference(self, depth_host):
        kp = self.params.kinfuParams
        depth = self.uploadDepth(depth_host)
        compute_dists(depth, kp.intr, out=self._dists)  # :55
        if self.frame_counter == 0 or self.warpfield is None:
            self.volume.integrate(self._dists, self.camera_pose, kp.intr)  # :70
            verts = self.volume.marchingCubes()[:, :3].contiguous()  # :74-85
            self.init(verts, verts)  # :88-95
            self.frame_counter += 1
            return False
        self.volume.clear()  # :114
        self.volume.integrate(self._dists, self.camera_pose, kp.intr)  # :115
        live = self.volume.marchingCubes()[:, :3].contiguous()  # :120-134
        self.liveVertices, self.liveNormals = live, live
        self.warpCanonicalToLiveOpt(live, paired=False)  # :140
        self.warpfield.update(self.canonicalWarpedToLive, self.params.blend_mode)  # :142
        self.frame_counter += 1
        return True

    # ---- one hot-path frame on device-resident inputs, integration overlapped with the next frame's point pipeline --
    # The warped integration of frame i needs only the node transforms solve(i) wrote; the point pipeline of frame i+1
    # (compute_dists, warp of the canonical frame, 8-NN graph, matrix pattern: ~0.1 ms of small kernels that leave most
    # SMs idle) needs neither the volume nor anything the integrator writes.  So integrate(i) runs on a second stream and
    # the only ordering added is: solve(i+1), which overwrites the transforms, waits for integrate(i).  Two dists buffers
    # alternate.  Results are bit-identical to the sequential schedule (tests/test_gpu_frontend.py).
    def _overlap_state(self):
        st = getattr(self, "_ov", None)
        if st is None:
            st = dict(stream=torch.cuda.Stream(self.device), dists=[self._dists, torch.empty_like(self._dists)], slot=0,
                      solved=torch.cuda.Event(), int_done=torch.cuda.Event(), pending=False)
            self._ov = st
        return st

    def frameDevice(self, depth_dev, liveVertices, overlap=True, timers=None):
        """compute_dists -> warpToLive(canonical) -> solve -> warped integration of `depth_dev` (int16/uint16 CUDA tensor).
        timers: optional dict of lists; CUDA event pairs around the solve ('solve') and the integration ('integrate')."""
        kp = self.params.kinfuParams
        cur = torch.cuda.current_stream(self.device)
        st = self._overlap_state()
        d = st["dists"][st["slot"]] if overlap else self._dists
        st["slot"] ^= 1 if overlap else 0
        compute_dists(depth_dev, kp.intr, out=d)
        self.canonicalWarpedToLive, _ = self.warpCanonical()
        self.solver.initializeProblemInstance(self.canonicalWarpedToLive, liveVertices)
        if st["pending"]:
            cur.wait_event(st["int_done"])  # the integration of the previous frame still reads the transforms
            st["pending"] = False
        if timers is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(cur)
        self.solver.solveAll()
        if timers is not None:
            b.record(cur)
            timers.setdefault("solve", []).append((a, b))
        istream = st["stream"] if overlap else cur
        if overlap:
            st["solved"].record(cur)
            istream.wait_event(st["solved"])
        with torch.cuda.stream(istream):
            if timers is not None:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(istream)
            self.volume.integrate(d, self.camera_pose, kp.intr, self.warpfield, self.params.blend_mode)
            if timers is not None:
                b.record(istream)
                timers.setdefault("integrate", []).append((a, b))
            if overlap:
                st["int_done"].record(istream)
                st["pending"] = True
        self.frame_counter += 1

    def frameSync(self):
        """make the current stream wait for an integration still running on the overlap stream"""
        st = getattr(self, "_ov", None)
        if st is not None and st["pending"]:
            torch.cuda.current_stream(self.device).wait_event(st["int_done"])
            st["pending"] = False

    # ---- pipelined frame loop for a steady stream of sensor frames -------------------------------------------------
    # streamFrame(depth, live) uploads this frame's pinned host inputs on a copy stream, runs the frame on the current
    # stream, brings {node transforms, solver statistics} back to pinned host memory on a second copy stream, and returns
    # the result of the PREVIOUS frame: the transfers of frame i overlap the kernels of frames i-1 / i+1 (two staging
    # slots), every frame still pays its own H2D and D2H.
    def _stream_state(self, live_host):
        st = getattr(self, "_ss", None)
        n = self.warpfield.numNodes()
        if st is not None and (st["live"][0].shape != live_host.shape or st["dq_dev"][0].shape[0] != n):
            self.streamFlush()  # sizes changed (new frame size, or Warpfield::update added nodes): start over
            torch.cuda.synchronize(self.device)
            st = None
        if st is None:
            st = dict(h2d=torch.cuda.Stream(self.device), d2h=torch.cuda.Stream(self.device), slot=0, pending=None,
                      depth=[torch.empty_like(self._depth_dev) for _ in range(2)],
                      live=[torch.empty(live_host.shape, dtype=torch.float32, device=self.device) for _ in range(2)],
                      dq_dev=[torch.empty((n, 8), dtype=torch.float32, device=self.device) for _ in range(2)],
                      stats_dev=[torch.empty(4, dtype=torch.float64, device=self.device) for _ in range(2)],
                      dq_host=[torch.empty((n, 8), dtype=torch.float32).pin_memory() for _ in range(2)],
                      stats_host=[torch.empty(4, dtype=torch.float64).pin_memory() for _ in range(2)],
                      up=[torch.cuda.Event() for _ in range(2)], free=[torch.cuda.Event() for _ in range(2)],
                      done=[torch.cuda.Event() for _ in range(2)], down=[torch.cuda.Event() for _ in range(2)])
            for e in st["free"] + st["down"]:
                e.record(torch.cuda.current_stream(self.device))
            self._ss = st
        return st

    def streamFrame(self, depth_host, live_host):
        """depth_host: pinned uint16/int16 [rows, cols]; live_host: pinned float32 [P, 3] paired with the canonical vertices.
        Returns None on the first call, afterwards (node transforms [N, 8], {energies, iterations}) of the previous frame;
        the transforms are a pinned staging buffer that is overwritten two calls later."""
        from ._lib import check, dptr, lib, stream_ptr
        kp = self.params.kinfuParams
        st = self._stream_state(live_host)
        s = st["slot"]
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(st["h2d"]):
            st["h2d"].wait_event(st["free"][s])  # the frame that last used this slot has consumed it
            st["depth"][s].copy_(depth_host.view(torch.int16) if depth_host.dtype == torch.uint16 else depth_host, non_blocking=True)
            st["live"][s].copy_(live_host, non_blocking=True)
            st["up"][s].record(st["h2d"])
        cur.wait_event(st["up"][s])
        cur.wait_event(st["down"][s])  # the result buffers of this slot have been read back
        self.frameDevice(st["depth"][s], st["live"][s], overlap=self.stream_overlap)
        self.frame_counter -= 1  # (counted below)
        check(lib.dfu_warpfield_get_nodes(self.warpfield.handle, None, dptr(st["dq_dev"][s]), None, stream_ptr(device=self.device)))
        self.solver.getStatsAsync(st["stats_dev"][s])
        st["free"][s].record(cur)
        st["done"][s].record(cur)
        with torch.cuda.stream(st["d2h"]):
            st["d2h"].wait_event(st["done"][s])
            st["dq_host"][s].copy_(st["dq_dev"][s], non_blocking=True)
            st["stats_host"][s].copy_(st["stats_dev"][s], non_blocking=True)
            st["down"][s].record(st["d2h"])
        self.frame_counter += 1
        prev, st["pending"], st["slot"] = st["pending"], s, 1 - s
        if prev is None:
            return None
        st["down"][prev].synchronize()
        e = st["stats_host"][prev]
        return st["dq_host"][prev], dict(initial_energy=float(e[0]), final_energy=float(e[1]), pcg_iterations=int(e[2]),
                                         gn_steps=int(e[3]))

    def streamFlush(self):
        """result of the last streamFrame call (waits for it, and for an integration still running on the overlap stream)"""
        self.frameSync()
        st = getattr(self, "_ss", None)
        if st is None or st["pending"] is None:
            return None
        p = st["pending"]
        st["pending"] = None
        st["down"][p].synchronize()
        e = st["stats_host"][p]
        return st["dq_host"][p], dict(initial_energy=float(e[0]), final_energy=float(e[1]), pcg_iterations=int(e[2]),
                                      gn_steps=int(e[3]))

    # DynFusion::operator() (src/dynfu/dyn_fusion.cpp:48-145), hot-path steps
    def __call__(self, depth_host, liveVertices=None):
        kp = self.params.kinfuParams
        depth = self.uploadDepth(depth_host)
        compute_dists(depth, kp.intr, out=self._dists)  # :55
        if self.frame_counter == 0 or self.warpfield is None:
            self.volume.integrate(self._dists, self.camera_pose, kp.intr)  # :70 (rigid, canonical frame)
        else:
            if liveVertices is not None:
                self.warpCanonicalToLiveOpt(liveVertices)  # :140
            # non-rigid surface fusion (README.md:14-17 "future additions"): live depth into the canonical volume
            self.volume.integrate(self._dists, self.camera_pose, kp.intr, self.warpfield, self.params.blend_mode)
        self.frame_counter += 1
        return True
