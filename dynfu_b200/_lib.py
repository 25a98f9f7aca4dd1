"""ctypes loader for libdynfu_b200.so (the C-ABI of include/dynfu_b200.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(_HERE, "libdynfu_b200.so")

BLEND_REF_COMPOSE = 0
BLEND_DQB_SUM = 1
NORMAL_REF = 0
NORMAL_ROTATE_ONLY = 1

STATUS = {0: "DFU_OK", 1: "DFU_ERR_INVALID", 2: "DFU_ERR_CUDA", 3: "DFU_ERR_PRECONDITION", 4: "DFU_ERR_NOT_INIT"}


class DfuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (STATUS.get(code, code), msg))
        self.code = code


class FrameParams(C.Structure):
    """dfu_frame_params (include/dynfu_b200.h)"""
    _fields_ = [("volume", C.c_void_p), ("dims", C.c_int * 3), ("voxel_size", C.c_float * 3), ("trunc_dist", C.c_float),
                ("max_weight", C.c_int), ("vol2cam", C.c_float * 12), ("intr", C.c_float * 4), ("rows", C.c_int), ("cols", C.c_int),
                ("blend_mode", C.c_int), ("z0", C.c_int), ("z1", C.c_int)]


class SolverParams(C.Structure):
    _fields_ = [
        ("num_iter", C.c_int),
        ("nonlinear_iter", C.c_int),
        ("linear_iter", C.c_int),
        ("tukey_offset", C.c_float),
        ("psi_data", C.c_float),
        ("lambda_", C.c_float),
        ("psi_reg", C.c_float),
        ("pcg_tol", C.c_float),
        ("early_out", C.c_int),
    ]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p)

if not os.path.exists(lib_path):
    raise ImportError(
        "dynfu_b200: %s is missing -- build it with `python -m dynfu_b200.build` (or __graft_entry__.build()). "
        "There is no CPU fallback." % lib_path)

lib = C.CDLL(lib_path)

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_SIGS = {
    "dfu_version": ([], C.c_int),
    "dfu_last_error": ([], C.c_char_p),
    "dfu_launch_count": ([], C.c_ulonglong),
    "dfu_device_check": ([_i], _i),
    "dfu_microbench_fp32": ([_i, C.POINTER(C.c_double), C.POINTER(C.c_double)], _i),
    "dfu_warpfield_create": ([C.POINTER(_vp), _i], _i),
    "dfu_warpfield_destroy": ([_vp], _i),
    "dfu_warpfield_init": ([_vp, _f, _vp, _vp, _vp, _i, _vp], _i),
    "dfu_warpfield_init_host": ([_vp, _f, _vp, _vp, _vp, _i, _vp], _i),
    "dfu_warpfield_num_nodes": ([_vp, C.POINTER(_i)], _i),
    "dfu_warpfield_get_nodes": ([_vp, _vp, _vp, _vp, _vp], _i),
    "dfu_warpfield_get_nodes_host": ([_vp, _vp, _vp, _vp, _vp], _i),
    "dfu_warpfield_set_transforms": ([_vp, _vp, _vp], _i),
    "dfu_warpfield_set_transforms_host": ([_vp, _vp, _vp], _i),
    "dfu_warpfield_update_translations": ([_vp, _vp, _vp], _i),
    "dfu_warpfield_knn": ([_vp, _vp, _i, _vp, _vp, _vp], _i),
    "dfu_warpfield_blend": ([_vp, _vp, _i, _vp, _i, _vp], _i),
    "dfu_warpfield_warp": ([_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _vp], _i),
    "dfu_compute_dists": ([_vp, _sz, _vp, _sz, _i, _i, C.POINTER(_f), _vp], _i),
    "dfu_tsdf_trunc_dist": ([_f, C.POINTER(_f)], _f),
    "dfu_tsdf_clear": ([_vp, C.POINTER(_i), _i, _i, _vp], _i),
    "dfu_tsdf_integrate": ([_vp, C.POINTER(_i), C.POINTER(_f), _f, _i, C.POINTER(_f), C.POINTER(_f), _vp, _sz, _i, _i,
                            _vp, _i, _i, _i, _vp], _i),
    "dfu_frame": ([_vp, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp, C.c_ulonglong, _vp, _vp, _i, _vp], _i),
    "dfu_marching_cubes": ([_vp, C.POINTER(_i), C.POINTER(_f), _vp, _vp, C.c_long, _vp, _vp], _i),
    "dfu_tsdf_integrate_stats": ([C.POINTER(C.c_ulonglong), _vp], _i),
    "dfu_solver_create": ([C.POINTER(_vp), _vp, C.POINTER(SolverParams)], _i),
    "dfu_solver_destroy": ([_vp], _i),
    "dfu_solver_set_allreduce": ([_vp, ALLREDUCE_FN, _vp], _i),
    "dfu_comm_unique_id": ([C.c_char_p], _i),
    "dfu_comm_create": ([C.POINTER(_vp), C.c_char_p, _i, _i], _i),
    "dfu_comm_destroy": ([_vp], _i),
    "dfu_comm_allreduce": ([_vp, _sz, _vp, _vp], _i),
    "dfu_solver_set_comm": ([_vp, _vp], _i),
    "dfu_solver_init_problem": ([_vp, _vp, _vp, _vp, _vp, _i, C.POINTER(_f), _vp], _i),
    "dfu_solver_solve_all": ([_vp, _vp], _i),
    "dfu_solver_huber_weights": ([_vp, _vp, _vp], _i),
    "dfu_solver_tukey_weights": ([_vp, _vp, _vp], _i),
    "dfu_solver_get_translations": ([_vp, _vp, _vp], _i),
    "dfu_solver_get_stats_host": ([_vp, C.POINTER(C.c_double), _vp], _i),
    "dfu_solver_get_stats": ([_vp, _vp, _vp], _i),
    "dfu_solver_set_energy": ([_vp, _i], _i),
    "dfu_solver_set_regulariser": ([_vp, _i], _i),
    "dfu_solver_get_increments": ([_vp, _vp, _vp], _i),
    "dfu_compute_points_normals": ([_vp, _sz, _i, _i, C.POINTER(_f), _vp, _sz, _vp, _sz, _vp], _i),
    "dfu_compact_points": ([_vp, _sz, _vp, _sz, _i, _i, C.POINTER(_f), _vp, _vp, _i, _vp, _vp], _i),
    "dfu_warpfield_unsupported": ([_vp, _vp, _i, _vp, _vp], _i),
    "dfu_voxel_grid_filter": ([_vp, _i, _f, _vp, C.POINTER(_i), _vp], _i),
    "dfu_warpfield_update": ([_vp, _vp, _i, _i, C.POINTER(_i), C.POINTER(_i), _vp], _i),
    "dfu_warpfield_cache_stats": ([_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), _vp], _i),
    "dfu_tsdf_raycast": ([_vp, C.POINTER(_i), C.POINTER(_f), _f, C.POINTER(_f), C.POINTER(_f), C.POINTER(_f), _i, _i, _f, _f, _vp, _sz,
                         _vp, _sz, _vp, _sz, _vp], _i),
    "dfu_pointcache_create": ([C.POINTER(_vp), _i], _i),
    "dfu_pointcache_destroy": ([_vp], _i),
    "dfu_warpfield_warp_cached": ([_vp, _vp, C.c_ulonglong, _vp, _vp, _i, _vp, _vp, _i, _i, _vp], _i),
    "dfu_pointindex_create": ([C.POINTER(_vp), _i], _i),
    "dfu_pointindex_destroy": ([_vp], _i),
    "dfu_pointindex_build": ([_vp, _vp, _i, _vp], _i),
    "dfu_pointindex_nearest": ([_vp, _vp, _i, _vp, _vp, _vp], _i),
    "dfu_find_corresponding": ([_vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp], _i),
}
EXPORTS = sorted(_SIGS)
for _name, (_args, _res) in _SIGS.items():
    _fn = getattr(lib, _name)  # raises AttributeError if the library does not export a declared symbol
    _fn.argtypes = _args
    _fn.restype = _res


def check(rc):
    if rc != 0:
        raise DfuError(rc, (lib.dfu_last_error() or b"").decode())


def farr(vals):
    return (C.c_float * len(vals))(*[float(v) for v in vals])


def iarr(vals):
    return (C.c_int * len(vals))(*[int(v) for v in vals])


def stream_ptr(stream=None, device=None):
    """cudaStream_t of `stream`, or of torch's current stream ON `device` (the device that owns the object the call is made
    for -- not the caller's current device)."""
    import torch

    s = torch.cuda.current_stream(device) if stream is None else stream
    return C.c_void_p(s.cuda_stream)


def dptr(t):
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise DfuError(1, "expected a CUDA tensor: this package has no CPU path")
    if not t.is_contiguous():
        raise DfuError(1, "expected a contiguous tensor")
    return C.c_void_p(t.data_ptr())


def dptr2d(t):
    """Device pointer of a row-pitched image (rows may be strided, everything inside a row is dense)."""
    if not t.is_cuda:
        raise DfuError(1, "expected a CUDA tensor: this package has no CPU path")
    inner = 1
    for d in range(t.dim() - 1, 0, -1):
        if t.stride(d) != inner:
            raise DfuError(1, "expected dense rows")
        inner *= t.shape[d]
    return C.c_void_p(t.data_ptr())


def hptr(a):
    """Host pointer of a numpy array / CPU tensor."""
    if a is None:
        return C.c_void_p(0)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)
