"""CPU-side checks of the drop-in boundary: the library loads without a GPU, exports every symbol that
include/dynfu_b200.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from dynfu_b200 import build as b

    path = b.build()
    assert os.path.exists(path)
    return C.CDLL(path)


def _declared():
    src = open(os.path.join(ROOT, "include", "dynfu_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dfu_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(built_lib):
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(built_lib, n), "libdynfu_b200.so does not export %s" % n


def test_python_binding_covers_the_header(built_lib):
    from dynfu_b200 import _lib

    assert sorted(_lib.EXPORTS) == [n for n in _declared() if n != "dfu_allreduce_fn"]


def test_no_cpu_fallback(built_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    built_lib.dfu_last_error.restype = C.c_char_p
    h = C.c_void_p()
    rc = built_lib.dfu_warpfield_create(C.byref(h), 0)
    assert rc == 2 and built_lib.dfu_last_error()  # DFU_ERR_CUDA, with a message
    import dynfu_b200

    with pytest.raises(dynfu_b200.DfuError):
        dynfu_b200.Warpfield()


def test_product_never_touches_the_oracle():
    """The product path must not import, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "dynfu_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "dynfu_oracle" not in txt and "oracle/" not in txt, f
    import subprocess

    out = subprocess.run(["ldd", os.path.join(pkg, "libdynfu_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
