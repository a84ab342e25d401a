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
# GAP-folded pooling (real geometry, small)
Fm = [torch.relu(torch.randn(3, 5, 5, 2048, device=dev)).requires_grad_(True) for _ in range(6)]
out = model.pool_fuse_gap(Fm, torch.randint(0, 10, (3, 6), dtype=torch.int32, device=dev), 10)
out.sum().backward()
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
