"""Builds libgvcnn_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python gvcnn-tf_b200/build.py [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libgvcnn_sm100.so")
SOURCES = ["capi.cu", "host_pipeline.cu", "comm.cu", "score.cu", "gap_score.cu", "scheme.cu", "pool_fwd.cu", "pool_fwd_ring.cu", "pool_gap_ring.cu", "pool_bwd.cu", "pool_bwd_fast.cu", "paper_mode.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "ring_common.cuh"), os.path.join(ROOT, "include", "gvcnn_b200.h")]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libgvcnn_sm100.so cannot be built (there is no fallback path)")


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines / out: A/B builds of the same sources with -D flags into another path (scripts/_build/)."""
    if out is None and not force and not needs_build():
        return SO
    cmd = [
        nvcc_path(), "-std=c++17", "-O3", "-lineinfo",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-fmad=false",                      # one rounding per float32 op unless fmaf() is written
        "-Xcompiler", "-fPIC", "-shared",
        "-I", os.path.join(ROOT, "include"), "-I", CSRC,
        "-o", out or SO,
    ] + ["-D" + d for d in defines] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    env = dict(os.environ)
    # the image's $CC/$CXX wrappers are not what nvcc should use as host compiler
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libgvcnn_sm100.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return out or SO


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, out=outs[0] if outs else None))
