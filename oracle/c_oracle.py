"""ctypes wrapper over oracle/libgvcnn_oracle.so.  TEST INFRASTRUCTURE ONLY
(see the header of oracle/gvcnn_oracle.c)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgvcnn_oracle.so")
_lib = None

POOL = {"max": 0, "mean": 1}


def build(force=False):
    src = os.path.join(_HERE, "gvcnn_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libgvcnn_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.oracle_score_f32.restype = ctypes.c_float
        L.oracle_score_f32.argtypes = [ctypes.c_float]
        L.oracle_score_f32_literal.restype = ctypes.c_float
        L.oracle_score_f32_literal.argtypes = [ctypes.c_float]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _strides(layout, B, V, D):
    """element strides (batch, view) of a [B,V,D] ('bvd') or [V,B,D] ('vbd') array"""
    if layout == "bvd":
        return V * D, D
    if layout == "vbd":
        return D, B * D
    raise ValueError(layout)


def view_score_x_f64(R, W, b, layout="bvd"):
    R = np.ascontiguousarray(R, dtype=np.float32)
    B, V, C = R.shape if layout == "bvd" else (R.shape[1], R.shape[0], R.shape[2])
    W = np.ascontiguousarray(W, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    x = np.empty((B, V), dtype=np.float64)
    sb, sv = _strides(layout, B, V, C)
    lib().oracle_view_score_x_f64(_p(R), _p(W), _p(b), _p(x), B, V, C,
                                  ctypes.c_int64(sb), ctypes.c_int64(sv))
    return x


def view_score_x_kernel_order(R, W, b, E=4, layout="bvd", lanes=None):
    """x in the CUDA score kernels' summation order.  ``lanes`` = lanes per row (32 in both the
    generic and the 4-rows-per-warp kernel of csrc/score.cu)."""
    R = np.ascontiguousarray(R, dtype=np.float32)
    B, V, C = R.shape if layout == "bvd" else (R.shape[1], R.shape[0], R.shape[2])
    if lanes is None:
        lanes = 32
    W = np.ascontiguousarray(W, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    x = np.empty((B, V), dtype=np.float32)
    sb, sv = _strides(layout, B, V, C)
    lib().oracle_view_score_x_kernel_order(_p(R), _p(W), _p(b), _p(x), B, V, C,
                                           ctypes.c_int64(sb), ctypes.c_int64(sv), E, lanes)
    return x


def score_f32(x):
    L = lib()
    x = np.asarray(x, dtype=np.float32)
    return np.array([L.oracle_score_f32(float(v)) for v in x.reshape(-1)],
                    dtype=np.float32).reshape(x.shape)


def score_f32_literal(x):
    L = lib()
    x = np.asarray(x, dtype=np.float32)
    return np.array([L.oracle_score_f32_literal(float(v)) for v in x.reshape(-1)],
                    dtype=np.float32).reshape(x.shape)


def bins(s, G):
    s = np.ascontiguousarray(s, dtype=np.float32)
    out = np.empty(s.shape, dtype=np.int32)
    lib().oracle_bins(_p(s), _p(out), ctypes.c_int64(s.size), G)
    return out


def pool_fuse_fwd(F, bins_, G, pool="max", empty_fill=1.0, layout="bvd", want_mask=False):
    F = np.ascontiguousarray(F, dtype=np.float32)
    B, V, D = F.shape if layout == "bvd" else (F.shape[1], F.shape[0], F.shape[2])
    bins_ = np.ascontiguousarray(bins_, dtype=np.int32)
    bin_sb = 0 if bins_.ndim == 1 else V
    S = np.empty((B, D), dtype=np.float32)
    mask = np.zeros(((V + 7) // 8, B, D), dtype=np.uint8) if want_mask else None
    sb, sv = _strides(layout, B, V, D)
    rc = lib().oracle_pool_fuse_fwd(_p(F), _p(bins_), _p(S), _p(mask) if want_mask else None,
                                    B, V, D, G, POOL[pool], ctypes.c_float(empty_fill),
                                    ctypes.c_int64(sb), ctypes.c_int64(sv), ctypes.c_int64(bin_sb))
    if rc != 0:
        raise IndexError("bin outside [0, %d)" % G)
    return (S, mask) if want_mask else S


def pool_fuse_bwd(dS, F, bins_, G, pool="max", layout="bvd"):
    F = np.ascontiguousarray(F, dtype=np.float32)
    B, V, D = F.shape if layout == "bvd" else (F.shape[1], F.shape[0], F.shape[2])
    dS = np.ascontiguousarray(dS, dtype=np.float32)
    bins_ = np.ascontiguousarray(bins_, dtype=np.int32)
    bin_sb = 0 if bins_.ndim == 1 else V
    dF = np.zeros_like(F)
    sb, sv = _strides(layout, B, V, D)
    rc = lib().oracle_pool_fuse_bwd(_p(dS), _p(F), _p(bins_), _p(dF), B, V, D, G, POOL[pool],
                                    ctypes.c_int64(sb), ctypes.c_int64(sv),
                                    ctypes.c_int64(sb), ctypes.c_int64(sv), ctypes.c_int64(bin_sb))
    if rc != 0:
        raise IndexError("bin outside [0, %d)" % G)
    return dF
