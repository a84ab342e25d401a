"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/gvcnn_b200.h declares, rejects bad arguments before launching anything, and the Python
mirror refuses to run without CUDA (no fallback).  No GPU compute is attempted here."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cabi():
    from gvcnn_tf_b200 import _cabi
    import importlib.util
    spec = importlib.util.spec_from_file_location("gvcnn_build", os.path.join(ROOT, "gvcnn-tf_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build()
    return _cabi


def declared_symbols():
    with open(os.path.join(ROOT, "include", "gvcnn_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(gvcnn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(cabi):
    syms = declared_symbols()
    assert len(syms) >= 16
    L = cabi.lib()
    for s in syms:
        assert hasattr(L, s), "declared in include/gvcnn_b200.h but not exported: " + s
        assert s in cabi.SIGNATURES, "exported but not bound in _cabi.py: " + s
    assert sorted(cabi.SIGNATURES) == syms
    assert L.gvcnn_version() == 2
    assert b"ok" == L.gvcnn_strerror(0)
    for code in range(-9, 0):
        assert b"unknown" not in L.gvcnn_strerror(code)


def test_header_constants_match_python(cabi):
    with open(os.path.join(ROOT, "include", "gvcnn_b200.h")) as f:
        text = f.read()
    defs = dict(re.findall(r"#define\s+(GVCNN_[A-Z0-9_]+)\s+\(?(-?\d+)\)?", text))
    assert int(defs["GVCNN_F32"]) == cabi.F32 and int(defs["GVCNN_BF16"]) == cabi.BF16
    assert int(defs["GVCNN_LAYOUT_BVD"]) == cabi.LAYOUT_BVD and int(defs["GVCNN_LAYOUT_PTRS"]) == cabi.LAYOUT_PTRS
    assert int(defs["GVCNN_POOL_MAX"]) == cabi.POOL_MAX and int(defs["GVCNN_POOL_MEAN"]) == cabi.POOL_MEAN
    assert int(defs["GVCNN_STATUS_WORDS"]) == cabi.STATUS_WORDS
    assert int(defs["GVCNN_STATUS_BAD_SCHEME"]) == cabi.STATUS_BAD_SCHEME
    assert int(defs["GVCNN_MAX_VIEWS"]) == cabi.MAX_VIEWS and int(defs["GVCNN_MAX_GROUPS"]) == cabi.MAX_GROUPS
    assert int(defs["GVCNN_FLAG_NEAR_EDGE"]) == cabi.FLAG_NEAR_EDGE


def test_argument_errors_before_any_launch(cabi):
    L = cabi.lib()
    buf = (ctypes.c_char * 4096)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    # null pointers / bad dims / bad enums are rejected with negative codes, nothing is launched
    assert L.gvcnn_pool_fuse_fwd(None, p, 12, None, 0, p, None, None, None, 4, 12, 64, 8, 0, 1.0, 0, 0, None) == -1
    assert L.gvcnn_pool_fuse_fwd(p, p, 12, None, 0, p, None, None, None, 0, 12, 64, 8, 0, 1.0, 0, 0, None) == 0    # empty batch: no-op
    assert L.gvcnn_pool_fuse_fwd(p, p, 12, None, 0, p, None, None, None, -1, 12, 64, 8, 0, 1.0, 0, 0, None) == -1
    assert L.gvcnn_pool_fuse_fwd(p, p, 12, None, 0, p, None, None, None, 4, 0, 64, 8, 0, 1.0, 0, 0, None) == -1
    assert L.gvcnn_pool_fuse_fwd(p, p, 12, None, 0, p, None, None, None, 4, 129, 64, 8, 0, 1.0, 0, 0, None) == -4
    assert L.gvcnn_pool_fuse_fwd(p, p, 12, None, 0, p, None, None, None, 4, 12, 64, 5000, 0, 1.0, 0, 0, None) == -5
    assert L.gvcnn_pool_fuse_fwd(p, p, 12, None, 0, p, None, None, None, 4, 12, 64, 8, 0, 1.0, 0, 7, None) == -2
    assert L.gvcnn_pool_fuse_fwd(p, p, 12, None, 0, p, None, None, None, 4, 12, 64, 8, 0, 1.0, 9, 0, None) == -3
    assert L.gvcnn_pool_fuse_fwd(p, p, 12, None, 0, p, None, None, None, 4, 12, 64, 8, 3, 1.0, 0, 0, None) == -8
    assert L.gvcnn_pool_fuse_bwd(p, p, 12, None, 0, None, p, None, 4, 12, 64, 8, 0, 0, 0, None) == -1   # max needs mask
    odd = ctypes.c_void_p(p.value + 2)
    assert L.gvcnn_pool_fuse_fwd(odd, p, 12, None, 0, p, None, None, None, 4, 12, 64, 8, 0, 1.0, 0, 0, None) == -6
    assert L.gvcnn_score_bin_fwd(p, None, p, None, p, p, None, None, 4, 12, 64, 8, 0, 0, 1, 0, None) == -1
    # the A/B kernel-variant selector rides in bits 8..11 of `pool` (stateless); values above 3 are rejected
    assert L.gvcnn_pool_fuse_fwd(p, p, 12, None, 0, p, None, None, None, 4, 12, 64, 8, cabi.pool_variant(0, 7), 1.0, 0, 0, None) == -8
    assert L.gvcnn_host_workspace_bytes(4096, 256, 12, 1024, 2048, 0, 0, 0) > 256 * 12 * 2048 * 4 * 3
    assert (L.gvcnn_host_workspace_bytes(4096, 256, 12, 1024, 2048, 0, 0, 1)
            >= L.gvcnn_host_workspace_bytes(4096, 256, 12, 1024, 2048, 0, 0, 0) + 4096 * 12 * 4)
    assert L.gvcnn_host_workspace_bytes(4096, 0, 12, 1024, 2048, 0, 0, 0) == 0
    # the host entry point needs a pipeline object and validates its modes before touching the device
    assert L.gvcnn_grouping_fusion_host(None, p, p, p, p, p, None, None, None, None, None, 4, 12, 64, 64, 8, 0, 1.0,
                                        0, 0, 4, None, None, 2, p, 1 << 20) == -1
    assert L.gvcnn_grouping_fusion_host(None, p, p, p, p, p, None, None, None, None, None, 4, 12, 64, 64, 8, 0, 1.0,
                                        0, 5, 4, None, None, 2, p, 1 << 20) == -8
    assert L.gvcnn_score_bin(p, 1.0, None, p, p, None, None, 12, 8, -1, 0, 0, None, 0, None) == -1      # negative multiplier
    if not torch.cuda.is_available():
        assert L.gvcnn_check_device() == -7            # no CPU fallback: the library says so


def test_python_mirror_has_no_cpu_path():
    from gvcnn_tf_b200 import model
    F = torch.zeros(2, 3, 8)
    bins = torch.zeros(2, 3, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="no CPU path"):
        model.pool_fuse(F, bins, 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        model.score_bin(torch.zeros(2, 3, 8), torch.zeros(3, 8), torch.zeros(3), 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        model.view_pooling([torch.zeros(2, 8)] * 3, torch.zeros(4, 3, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="no CPU path"):
        model.group_fusion({0: torch.zeros(2)}, torch.ones(1))        # a plain dict is accepted, CPU tensors are not
    with pytest.raises(TypeError):
        model.group_fusion([torch.zeros(2)], torch.ones(1))


def test_missing_library_fails_loudly(monkeypatch):
    from gvcnn_tf_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "SO_PATH", "/nonexistent/libgvcnn_sm100.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _cabi.lib()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under gvcnn-tf_b200/ may reference it."""
    pkg = os.path.join(ROOT, "gvcnn-tf_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "libgvcnn_oracle" not in src, fn


def test_no_packed_fma_in_the_library(cabi):
    """ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even under -fmad=false, which breaks the
    one-rounding-per-op contract the bit-exact parity rests on.  The kernels therefore use packed ADDS only
    (FADD2, csrc/common.cuh); no FFMA2 may appear in the built library, while FADD2 must (the packed path is live)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", cabi.SO_PATH], capture_output=True, text=True, timeout=600).stdout
    assert sass.count("FFMA2") == 0
    assert sass.count("FADD2") > 100
