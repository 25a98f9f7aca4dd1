"""Builds libdynfu_b200.so (hand-written CUDA for sm_100a + the C-ABI) in-tree with nvcc.

No torch.utils.cpp_extension, no JIT cache: the .so lands next to this file so it travels to the GPU box
with the gpurun snapshot.  -fmad=false: the parity-critical arithmetic is written with explicit intrinsics
and must never be contracted; FMA is used only where the source says __fmaf_rn."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdynfu_b200.so")
SOURCES = ["warpfield.cu", "tsdf.cu", "solver.cu", "comm.cu", "frontend.cu", "update.cu", "raycast.cu", "microbench.cu", "marching_cubes.cu", "frame.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++", "--expt-relaxed-constexpr",
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".inc"))]
    headers.append(os.path.join(HERE, "..", "include", "dynfu_b200.h"))
    headers.append(os.path.abspath(__file__))
    objs = []

    def compile_one(src):
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        if force or _stale(obj, [os.path.join(CSRC, src)] + headers):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
            if verbose:
                print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(compile_one, srcs))
    if force or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-ccbin", "/usr/bin/g++", "-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return OUT


def _build_cpp(name):
    root = os.path.dirname(HERE)
    src = os.path.join(root, "tests", "cpp", name + ".cpp")
    out = os.path.join(root, "tests", "cpp", name)
    deps = [src, os.path.join(root, "tests", "cpp", "mini_gtest.h"), os.path.join(HERE, "adapter", "dynfu_adapter.hpp"), OUT]
    if _stale(out, deps):
        cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-I/usr/local/cuda/include", "-o", out, src, "-L" + HERE, "-ldynfu_b200",
               "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,$ORIGIN/../../dynfu_b200", "-Wl,-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building tests/cpp/%s failed:\n%s\n%s" % (name, r.stdout, r.stderr))
    return out


def build_cpp_tests(verbose=False):
    """tests/cpp/opt_test: the reference's gtest cases compiled against the adapter header + the C-ABI library
    (and tests/cpp/frame_test: the frame operator through the same header)."""
    _build_cpp("frame_test")
    return _build_cpp("opt_test")


def build_cpp_frame_test():
    return _build_cpp("frame_test")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_cpp_tests())
