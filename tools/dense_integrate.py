import sys, os, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import dynfu_b200 as dfu
from dynfu_b200._lib import lib
import bench
from tools import synth
scene = bench.make_scene()
devs = torch.device('cuda', 0)
def dev(a, dt=torch.float32): return torch.as_tensor(np.ascontiguousarray(a)).to(devs, dtype=dt)
prm = dfu.DynFuParams(kinfuParams=dfu.KinFuParams(volume_dims=(512,)*3), epsilon=bench.EPSILON, lambda_=bench.LAMBDA,
      solver=dfu.CombinedSolverParameters(numIter=5, nonLinearIter=1, linearIter=10, earlyOut=False, pcgTolerance=0.0))
dense = len(sys.argv) < 2 or sys.argv[1] != 'sparse'
W = synth.with_wall if dense else (lambda d: d)
df = dfu.DynFusion(prm, device=devs)
df.init(dev(scene["canon"]), None, nodes=(dev(scene["pos"]), dev(scene["dq"]), dev(scene["dg_w"])))
df(torch.from_numpy(W(scene["depth0"]).view(np.int16)).pin_memory())
df.warpCanonicalToLiveOpt(dev(scene["lives"][0]))
kp = prm.kinfuParams
dd = [dfu.compute_dists(dev(W(d).view(np.int16), torch.int16), kp.intr) for d in scene["depths"]]
for i in range(8): df.volume.integrate(dd[i % 4], df.camera_pose, kp.intr, df.warpfield, prm.blend_mode)
torch.cuda.synchronize()
ts = []
for i in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); df.volume.integrate(dd[i % 4], df.camera_pose, kp.intr, df.warpfield, prm.blend_mode); b.record()
    torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
st = (C.c_ulonglong * 4)(); lib.dfu_tsdf_integrate_stats(st, None)
print('dense' if dense else 'sparse', 'integrate ms median %.4f min %.4f' % (np.median(ts), np.min(ts)), 'updated', st[0], 'quads', st[1], 'saturated-path voxels', st[2], 'warped bricks', st[3])
