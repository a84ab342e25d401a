#!/usr/bin/env python
"""Training step in paper mode (score-derived weights, gradient to the score FC) at BASELINE configs[1] size."""
import json, os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvcnn_tf_b200 import model  # noqa: E402
B, V, D, G, Cr = 4096, 12, 2048, 8, 1024
dev = "cuda:0"
F = torch.relu(torch.randn(B, V, D, device=dev)).requires_grad_(True)
R = torch.randn(B, V, Cr, device=dev)
W = ((torch.rand(V, Cr, device=dev) * 2 - 1) * 0.0765).requires_grad_(True)
b = torch.zeros(V, device=dev, requires_grad=True)
dS = torch.randn(B, D, device=dev)
def step():
    F.grad = W.grad = b.grad = None          # no gradient accumulation kernels in the timed loop
    S, *_ = model.grouping_fusion_paper(R, W, b, F, G)
    S.backward(dS)
for _ in range(3): step()
torch.cuda.synchronize()
ts = []
for _ in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
print(json.dumps({"paper_mode_train_step_us": statistics.median(ts), "shapes_per_s": B / (statistics.median(ts) * 1e-6)}))
