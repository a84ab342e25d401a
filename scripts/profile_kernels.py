#!/usr/bin/env python
"""Launches each kernel of the path a few times at BASELINE configs[1] size, for ncu.

    ncu --set full --clock-control none --import-source on -k regex:'pool_fuse|view_score|batch_mean|gap_score' \
        -s 14 -c 7 -o gpurun_out/prof python scripts/profile_kernels.py
Order per round (7 kernels): view_score (fused bin), pool_fuse_fwd (no mask), pool_fuse_fwd (mask), pool_fuse_bwd, then the
literal batch-mode forward = view_score (x only), batch_mean_bin_fused, pool_fuse_fwd (one shared bin row); --gap adds
gap_score (N=128, V=6, 10x10x1024 maps) as an 8th.
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvcnn_tf_b200 import _cabi as C  # noqa: E402

B, V, D, G, Cr = 4096, 12, 2048, 8, 1024
dtype = torch.bfloat16 if "--bf16" in sys.argv else torch.float32
dt = C.BF16 if dtype == torch.bfloat16 else C.F32
pool = C.POOL_MEAN if "--mean" in sys.argv else C.POOL_MAX
rounds = 4
dev = torch.device("cuda:0")
torch.manual_seed(0)
F = torch.randn(B, V, D, device=dev).to(dtype)
R = torch.randn(B, V, Cr, device=dev).to(dtype)
W = (torch.rand(V, Cr, device=dev) * 2 - 1) * (6.0 / (Cr + 1)) ** 0.5
b = torch.zeros(V, device=dev)
dS = torch.randn(B, D, device=dev).to(dtype)
scores = torch.empty(B, V, device=dev)
bins = torch.empty(B, V, dtype=torch.int32, device=dev)
status = torch.zeros(4, dtype=torch.int32, device=dev)
S = torch.empty(B, D, device=dev, dtype=dtype)
mask = torch.empty((V + 7) // 8, B, D, dtype=torch.uint8, device=dev)
dF = torch.empty(B, V, D, device=dev, dtype=dtype)
L = C.lib()
p = lambda t: ctypes.c_void_p(t.data_ptr())
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(rounds):
    C.check(L.gvcnn_score_bin_fwd(p(R), p(W), p(b), None, p(scores), p(bins), None, p(status), B, V, Cr, G,
                                  C.LAYOUT_BVD, dt, 0, 1, sp), "score")
    C.check(L.gvcnn_pool_fuse_fwd(p(F), p(bins), V, None, 0, p(S), None, None, p(status), B, V, D, G, pool,
                                  ctypes.c_float(1.0), C.LAYOUT_BVD, dt, sp), "fwd")
    C.check(L.gvcnn_pool_fuse_fwd(p(F), p(bins), V, None, 0, p(S), None, p(mask), p(status), B, V, D, G, pool,
                                  ctypes.c_float(1.0), C.LAYOUT_BVD, dt, sp), "fwd+mask")
    C.check(L.gvcnn_pool_fuse_bwd(p(dS), p(bins), V, None, 0, p(mask), p(dF), p(status), B, V, D, G, pool,
                                  C.LAYOUT_BVD, dt, sp), "bwd")
    xb = torch.empty(B, V, device=dev)
    xsum = torch.empty(V, device=dev)
    sc1 = torch.empty(V, device=dev)
    bins1 = torch.empty(V, dtype=torch.int32, device=dev)
    bl = (torch.rand(V, device=dev) * 8 - 4)
    C.check(L.gvcnn_grouping_fusion_batch_fwd(p(R), p(W), p(bl), p(F), p(xb), p(xsum), None, p(sc1), p(bins1), None, p(S), None,
                                              p(status), B, V, Cr, D, G, 0, pool, ctypes.c_float(1.0), C.LAYOUT_BVD,
                                              C.LAYOUT_BVD, dt, 0, 1, B, None, None, sp), "batch fwd")
    if "--gap" in sys.argv:
        N2, V2, HW = 128, 6, 100
        maps = torch.randn(N2, V2, HW, Cr, device=dev).to(dtype)
        W2 = W[:V2].contiguous()
        s2 = torch.empty(N2, V2, device=dev)
        b2 = torch.empty(N2, V2, dtype=torch.int32, device=dev)
        C.check(L.gvcnn_gap_score_bin_fwd(p(maps), p(W2), p(b), None, None, p(s2), p(b2), None, p(status), N2, V2, HW, Cr, 10,
                                          C.LAYOUT_BVD, dt, 1, 0, 1, sp), "gap score")
torch.cuda.synchronize()
print("profiled", rounds, "rounds; status", status.tolist())
