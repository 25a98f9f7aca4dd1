"""CombinedSolver -- host mirror of the reference's class (include/dynfu/utils/opt_solver.hpp:19-110)."""
import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import check, dptr, lib, stream_ptr


@dataclass
class CombinedSolverParameters:
    """The fields of Opt's CombinedSolverParameters the reference sets (src/dynfu/dyn_fusion.cpp:183-189,
    test/opt_optimisation_test.cpp:38-44)."""
    numIter: int = 24
    nonLinearIter: int = 16
    linearIter: int = 256
    useOpt: bool = True
    useOptLM: bool = False
    earlyOut: bool = True
    optDoublePrecision: bool = False
    pcgTolerance: float = 1e-6  # not in Opt's struct: relative stop of the PCG (r.z <= tol^2 * first r.z)


class CombinedSolver:
    # CombinedSolver::CombinedSolver (src/dynfu/utils/opt_solver.cpp:3-13)
    def __init__(self, warpfield, params, tukeyOffset, psi_data, lambda_, psi_reg):
        self.warpfield = warpfield  # shares the nodes, like the reference's by-value copy of shared_ptr<Node>s
        self.device = warpfield.device
        self.params = params
        p = _lib.SolverParams(params.numIter, params.nonLinearIter, params.linearIter, tukeyOffset, psi_data, lambda_,
                              psi_reg, params.pcgTolerance, 1 if params.earlyOut else 0)
        h = C.c_void_p()
        check(lib.dfu_solver_create(C.byref(h), warpfield.handle, C.byref(p)))
        self._h = h
        self._cb = None
        self._keep = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.dfu_solver_destroy(h)
            self._h = None

    def setAllReduce(self, fn):
        """fn(tensor) sums a float32 CUDA tensor over ranks in place (data-parallel point partitions)."""
        dev = self.warpfield.device

        def _cb(buf, count, _ctx, stream):
            try:
                t = _tensor_from_ptr(buf, count, dev)
                if stream:  # NULL = the legacy default stream, which is torch's default stream too
                    with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=dev)):
                        fn(t)
                else:
                    with torch.cuda.stream(torch.cuda.default_stream(dev)):
                        fn(t)
                return 0
            except Exception:  # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return 1

        self._cb = _lib.ALLREDUCE_FN(_cb)
        check(lib.dfu_solver_set_allreduce(self._h, self._cb, None))

    def setCommunicator(self, comm):
        """dynfu_b200.dist.Communicator: all-reduces issued by the library itself through NCCL"""
        self._comm = comm
        check(lib.dfu_solver_set_comm(self._h, comm.handle if comm is not None else None))

    # CombinedSolver::initializeProblemInstance (src/dynfu/utils/opt_solver.cpp:15-54)
    def initializeProblemInstance(self, canonicalVertices, liveVertices, canonicalNormals=None, liveNormals=None,
                                  affine=None):
        dev = self.warpfield.device
        cv = torch.as_tensor(canonicalVertices, dtype=torch.float32).to(dev).reshape(-1, 3).contiguous()
        lv = torch.as_tensor(liveVertices, dtype=torch.float32).to(dev).reshape(-1, 3).contiguous()
        if cv.shape != lv.shape:
            raise _lib.DfuError(1, "canonical and live frames must pair up vertex by vertex")
        ln = None
        if liveNormals is not None:
            ln = torch.as_tensor(liveNormals, dtype=torch.float32).to(dev).reshape(-1, 3).contiguous()
            if ln.shape != lv.shape:
                raise _lib.DfuError(1, "live normals must pair up with the live vertices")
        self._keep = (cv, lv, ln)  # the point-to-plane mode reads them during solveAll
        check(lib.dfu_solver_init_problem(self._h, dptr(cv), None, dptr(lv), dptr(ln), cv.shape[0], None, stream_ptr(device=self.device)))

    # north-star extension (no reference implementation): point-to-plane data term with a rigid increment per node
    ENERGY_REF_TRANSLATION = 0
    ENERGY_P2PLANE_SE3 = 1

    def setEnergy(self, mode):
        """call before initializeProblemInstance; ENERGY_P2PLANE_SE3 needs liveNormals there"""
        check(lib.dfu_solver_set_energy(self._h, int(mode)))

    REG_QUADRATIC = 0
    REG_HUBER_ALPHA = 1

    def setRegulariser(self, mode):
        """ENERGY_P2PLANE_SE3 only: REG_QUADRATIC (default) or REG_HUBER_ALPHA = DynamicFusion eq. 8, alpha_ij = max(dg_w)
        times the Huber function with threshold psi_reg (the term the reference prepares and leaves out,
        opt_solver.cpp:233-268, energy.t:76)"""
        check(lib.dfu_solver_set_regulariser(self._h, int(mode)))

    def getIncrements(self):
        """ENERGY_P2PLANE_SE3: [N, 12] rigid increments (R row-major, t) of the last solve"""
        x = torch.empty((self.warpfield.numNodes(), 12), dtype=torch.float32, device=self.warpfield.device)
        check(lib.dfu_solver_get_increments(self._h, dptr(x), stream_ptr(device=self.device)))
        return x

    # CombinedSolverBase::solveAll [Opt]
    def solveAll(self):
        check(lib.dfu_solver_solve_all(self._h, stream_ptr(device=self.device)))

    # CombinedSolver::updateHuberWeights (src/dynfu/utils/opt_solver.cpp:241-268)
    def huberWeights(self):
        h = torch.empty((self.warpfield.numNodes(),), dtype=torch.float32, device=self.warpfield.device)
        check(lib.dfu_solver_huber_weights(self._h, dptr(h), stream_ptr(device=self.device)))
        return h

    # the tukey biweights (src/dynfu/utils/opt_solver.cpp:204-231) the last solve ended with
    def tukeyWeights(self):
        t = torch.empty((self._keep[0].shape[0],), dtype=torch.float32, device=self.warpfield.device)
        check(lib.dfu_solver_tukey_weights(self._h, dptr(t), stream_ptr(device=self.device)))
        return t

    def getTranslations(self):
        n = self.warpfield.numNodes()
        t = torch.empty((n, 3), dtype=torch.float32, device=self.warpfield.device)
        check(lib.dfu_solver_get_translations(self._h, dptr(t), stream_ptr(device=self.device)))
        return t

    def getStatsAsync(self, out=None):
        """the same four numbers as a float64 CUDA tensor [E0, E, pcg iterations, gn steps]: stream-ordered, no host sync"""
        if out is None:
            out = torch.empty(4, dtype=torch.float64, device=self.warpfield.device)
        check(lib.dfu_solver_get_stats(self._h, dptr(out), stream_ptr(device=self.device)))
        return out

    def getStats(self):
        """{initial energy, final energy, PCG iterations, GN steps} of the last solveAll (synchronises)."""
        s = (C.c_double * 4)()
        check(lib.dfu_solver_get_stats_host(self._h, s, stream_ptr(device=self.device)))
        return dict(initial_energy=s[0], final_energy=s[1], pcg_iterations=int(s[2]), gn_steps=int(s[3]))


def _tensor_from_ptr(ptr, count, device):
    """float32 CUDA tensor aliasing `count` floats at device address `ptr` (no copy)."""

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                  "strides": None}
    return torch.as_tensor(h, device=device)
