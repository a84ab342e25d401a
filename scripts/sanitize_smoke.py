#!/usr/bin/env python
"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py
Sizes are tiny so the instrumented run stays short; shapes are chosen to hit partial tiles and every path."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvcnn_tf_b200 import model  # noqa: E402

torch.manual_seed(0)
dev = "cuda:0"
cases = [  # B, V, D, G, dtype, pool
    (7, 12, 2048, 8, torch.float32, "max"),    # ring + bwd_fast
    (5, 12, 3076, 8, torch.float32, "max"),    # ring, partial last tile
    (5, 6, 1024, 10, torch.bfloat16, "max"),   # packed bf16
    (4, 20, 2048, 16, torch.float32, "mean"),  # V = 20: 128-column tiles, two CTAs per SM
    (3, 80, 1032, 4, torch.float32, "max"),    # chunked + plane fix-up, generic bwd
    (3, 40, 2048, 3, torch.bfloat16, "max"),   # chunked bf16
    (4, 5, 100, 5, torch.float32, "max"),      # generic one-shot (V not templated), small D
    (2, 12, 7, 8, torch.float32, "mean"),      # scalar fallback (unaligned D)
    (900, 4, 2048, 4, torch.float32, "max"),   # ring, 1800 tiles: ~6 per CTA, ring slots re-used
    (6, 12, 2048, 8, torch.bfloat16, "mean"),  # bf16 mean walk four rows at a time (mixed-precision adds, reciprocal)
    (4, 20, 1024, 16, torch.bfloat16, "mean"), # the same on the narrow tiles, V = 20
    (5, 20, 2048, 16, torch.bfloat16, "max"),  # bf16 backward with three tie planes
]
for B, V, D, G, dt, pool in cases:
    F = torch.relu(torch.randn(B, V, D, device=dev)).to(dt).requires_grad_(True)
    R = torch.randn(B, V, 1024, device=dev).to(dt)
    W = (torch.rand(V, 1024, device=dev) * 2 - 1) * 0.0765
    b = torch.zeros(V, device=dev)
    S, sr = model.grouping_fusion(R, W, b, F, G, pool=pool)
    S.backward(torch.randn_like(S))
    desc = model.view_pooling([F.detach()[:, v] for v in range(V)], model.group_scheme(sr.scores[:1], G, V), pool=pool)
    _ = desc[0]
    _ = model.group_fusion(desc, torch.rand(G, device=dev) + 0.5)
    sb = model.score_bin(R, W, b + 1.0, G, score_reduce="batch")
    torch.cuda.synchronize()
    # round 2: the literal forward in one call (fused batch-mean kernel), custom weights with the ones dummy on the
    # ring's weights instantiation, a plain dict through group_fusion, the literal multiplier
    S1, sr1 = model.grouping_fusion(R, W, b + 1.0, F.detach(), G, pool=pool, score_reduce="batch", clamp=True)
    _ = model.pool_fuse(F.detach(), sr.bins.clamp(0, G - 1), G, pool=pool, group_weight=torch.rand(G, device=dev) + 0.5)
    _ = model.group_fusion({g: F.detach()[:, g % V] for g in range(min(G, 6))}, torch.rand(G, device=dev) + 0.5)
    _ = model.group_scheme(sr.scores[:1] * 0.5, max(G, 10), V, multiplier=10)
    torch.cuda.synchronize()
# round 2: GAP of the raw maps folded into the score kernel (fp32 and bf16, per-shape and batch, odd HW)
for dt in (torch.float32, torch.bfloat16):
    maps = torch.randn(3, 6, 7, 1024, device=dev).to(dt)
    W6 = (torch.rand(6, 1024, device=dev) * 2 - 1) * 0.0765
    b6 = torch.zeros(6, device=dev)
    _ = model.score_bin(maps, W6, b6, 10, clamp=True, check=False)
    _ = model.score_bin([maps[:, v].reshape(3, 7, 1, 1024) for v in range(6)], W6, b6 + 1.0, 10, score_reduce="batch", clamp=True)
    torch.cuda.synchronize()
# round 2: one-rank communicator (push to self, poll, scale) and the host-buffer entry point in both modes
import ctypes  # noqa: E402
from gvcnn_tf_b200 import _cabi as C  # noqa: E402
L = C.lib()
comm = ctypes.c_void_p()
handle = (ctypes.c_ubyte * C.COMM_HANDLE_BYTES)()
C.check(L.gvcnn_comm_create(ctypes.byref(comm), 0, 1, handle), "comm_create")
C.check(L.gvcnn_comm_connect(comm, bytes(handle)), "comm_connect")
x = torch.randn(12300, device=dev)
for _ in range(3):
    C.check(L.gvcnn_comm_allreduce_scaled_f32(comm, ctypes.c_void_p(x.data_ptr()), x.numel(), ctypes.c_float(0.5),
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "allreduce")
torch.cuda.synchronize()
C.check(L.gvcnn_comm_error(comm), "comm_error")
L.gvcnn_comm_destroy(comm)
Bh, Vh, Dh, Gh, Ch = 300, 12, 1024, 8, 256
Fh, Rh = torch.randn(Bh, Vh, Dh).pin_memory(), torch.randn(Bh, Vh, Ch).pin_memory()
Sh = torch.empty(Bh, Dh).pin_memory()
Wh, bh = (torch.rand(Vh, Ch, device=dev) - 0.5), torch.zeros(Vh, device=dev) + 1.0
pipe = ctypes.c_void_p()
C.check(L.gvcnn_host_pipeline_create(ctypes.byref(pipe), 2), "pipeline_create")
p = lambda t: ctypes.c_void_p(t.data_ptr())
for mode in (0, 1):
    nbytes = L.gvcnn_host_workspace_bytes(Bh, 128, Vh, Ch, Dh, C.F32, 0, mode)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    C.check(L.gvcnn_grouping_fusion_host(pipe, p(Rh), p(Fh), p(Wh), p(bh), p(Sh), None, None, None, None, None, Bh, Vh, Ch,
                                         Dh, Gh, C.POOL_MAX, ctypes.c_float(1.0), C.F32, mode, Bh, None, None, 128, p(ws),
                                         nbytes), "host")
L.gvcnn_host_pipeline_destroy(pipe)
# GAP-folded pooling (real geometry, small)
Fm = [torch.relu(torch.randn(3, 5, 5, 2048, device=dev)).requires_grad_(True) for _ in range(6)]
out = model.pool_fuse_gap(Fm, torch.randint(0, 10, (3, 6), dtype=torch.int32, device=dev), 10)
out.sum().backward()
for Vg in (4, 16, 20):  # the view counts added late in round 2 (one CTA per SM for 16 and 20)
    Fm = [torch.relu(torch.randn(2, 2, 3, 2048, device=dev)).to(torch.bfloat16).requires_grad_(True) for _ in range(Vg)]
    out = model.pool_fuse_gap(Fm, torch.randint(0, 8, (2, Vg), dtype=torch.int32, device=dev), 8, pool="mean", empty_fill=0.0)
    out.float().sum().backward()
torch.cuda.synchronize()
# paper mode
F = torch.relu(torch.randn(6, 12, 512, device=dev)).requires_grad_(True)
R = torch.randn(6, 12, 256, device=dev).requires_grad_(True)
W = ((torch.rand(12, 256, device=dev) * 2 - 1) * 0.15).requires_grad_(True)
b = torch.zeros(12, device=dev, requires_grad=True)
S, *_ = model.grouping_fusion_paper(R, W, b, F, 8)
S.sum().backward()
torch.cuda.synchronize()
print("sanitize smoke ok")
