"""dynfu_b200 -- B200-native (sm_100a) DynamicFusion hot path behind the reference's interface.

Host-side mirror of the reference classes on top of the C-ABI of include/dynfu_b200.h:
    Warpfield        (include/dynfu/warp_field.hpp)          -> dynfu_b200.Warpfield
    CombinedSolver   (include/dynfu/utils/opt_solver.hpp)    -> dynfu_b200.CombinedSolver
    cuda::TsdfVolume (include/kfusion/cuda/tsdf_volume.hpp)  -> dynfu_b200.TsdfVolume
    DynFusion / KinFu frame operator                         -> dynfu_b200.DynFusion
PyTorch is used for device memory, streams and torch.distributed only.  There is no CPU fallback:
importing the package without the compiled extension, or calling it without a CUDA device, fails loudly.
"""
from ._lib import DfuError, lib, lib_path, BLEND_REF_COMPOSE, BLEND_DQB_SUM, NORMAL_REF, NORMAL_ROTATE_ONLY  # noqa: F401
from .warpfield import Warpfield  # noqa: F401
from .tsdf_volume import TsdfVolume, compute_dists  # noqa: F401
from .solver import CombinedSolver, CombinedSolverParameters  # noqa: F401
from .dyn_fusion import DynFusion, DynFuParams, KinFuParams  # noqa: F401
from . import frontend  # noqa: F401
