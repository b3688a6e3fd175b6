"""The C-ABI libraries load without a GPU and export every symbol include/*.h declares."""
import ctypes as C
import re
from pathlib import Path

import pytest

from whittedstyle_raytracer_b200 import cabi

REPO = Path(__file__).resolve().parent.parent


def declared_functions(header):
    text = re.sub(r"/\*.*?\*/", "", (REPO / "include" / header).read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(wrt_[a-z0-9_]+)\s*\(", text)))


def test_host_library_exports_header_symbols():
    lib = cabi.load_host()
    names = declared_functions("wrt_host.h")
    assert sorted(cabi.HOST_SYMBOLS) == names
    for n in names:
        assert getattr(lib, n) is not None


def test_cuda_library_exports_header_symbols():
    lib = cabi.load_cuda()          # loads on a CPU-only box: libcudart is linked statically
    names = declared_functions("wrt_cuda.h")
    assert sorted(cabi.CUDA_SYMBOLS) == names
    for n in names:
        assert getattr(lib, n) is not None


def test_struct_layouts_match_headers():
    assert C.sizeof(cabi.WrtNode) == 32
    assert C.sizeof(cabi.WrtMaterial) == 48
    assert C.sizeof(cabi.WrtLight) == 80
    assert C.sizeof(cabi.WrtHit) == 60
    assert C.sizeof(cabi.WrtCamera) == 4 * (7 * 3 + 1 + 3)
    assert C.sizeof(cabi.WrtStats) == 8 * 14 + 16 + 24


def test_cuda_library_is_sm100a_only():
    """No multi-arch fat binary, no PTX JIT path to other GPUs: sm_100a SASS only."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", str(cabi.CUDA_LIB)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = cabi.load_cuda()
    h = C.c_void_p()
    assert lib.wrt_create(0, C.byref(h)) != 0
    assert b"no CPU fallback" in lib.wrt_last_error()
