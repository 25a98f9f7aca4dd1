"""Writes tests/golden/mc_tables.npz: the marching-cubes tables of the reference (src/kfusion/marching_cubes.cpp:66-354) as
compiled into oracle/_ref/libdynfu_ref_cuda.so (oracle/Makefile, target refcuda).  Run in the container that has
/root/reference; the fixture travels to the GPU box."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyoracle  # noqa: E402

e, t, n = pyoracle.RefCuda().mc_tables()
assert np.array_equal(n, (t >= 0).sum(1))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mc_tables.npz"), edge=e.astype(np.int32),
                    tri=t.astype(np.int8), numverts=n.astype(np.int32))
print("written")
