#!/usr/bin/env python
"""Latency of the path at the reference's own training shape (train.py:94-99: batch_size 4, 6 views,
block4 maps [N,10,10,2048] -> D = 204800, num_group 10, C_raw 1024): plain stream launches vs a CUDA
graph of the two launches (score+bin, pool+fuse).  Prints one JSON line."""
import ctypes
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvcnn_tf_b200 import _cabi as C  # noqa: E402

B, V, D, G, Cr = 4, 6, 10 * 10 * 2048, 10, 1024
dev = torch.device("cuda:0")
L = C.lib()
torch.manual_seed(0)
F = torch.randn(B, V, D, device=dev)
R = torch.randn(B, V, Cr, device=dev)
W = (torch.rand(V, Cr, device=dev) * 2 - 1) * (6.0 / (Cr + 1)) ** 0.5
b = torch.zeros(V, device=dev)
scores = torch.empty(B, V, device=dev)
bins = torch.empty(B, V, dtype=torch.int32, device=dev)
status = torch.zeros(4, dtype=torch.int32, device=dev)
S = torch.empty(B, D, device=dev)
p = lambda t: ctypes.c_void_p(t.data_ptr())


def step(stream):
    sp = ctypes.c_void_p(stream.cuda_stream)
    C.check(L.gvcnn_score_bin_fwd(p(R), p(W), p(b), None, p(scores), p(bins), None, p(status), B, V, Cr, G,
                                  C.LAYOUT_BVD, C.F32, 0, 1, sp), "score")
    C.check(L.gvcnn_pool_fuse_fwd(p(F), p(bins), V, None, 0, p(S), None, None, p(status), B, V, D, G, C.POOL_MAX,
                                  ctypes.c_float(1.0), C.LAYOUT_BVD, C.F32, sp), "pool")


def timeit(fn, n=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


cur = torch.cuda.current_stream()
t_plain = timeit(lambda: step(cur))
S_ref = S.clone()
side = torch.cuda.Stream()
side.wait_stream(cur)
with torch.cuda.stream(side):
    for _ in range(3):
        step(side)
side.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=side):
    step(side)
S.zero_()
g.replay()
torch.cuda.synchronize()
same = bool(torch.equal(S, S_ref))
t_graph = timeit(lambda: g.replay())
bytes_moved = B * (V * D * 4 + D * 4 + V * Cr * 4)
print(json.dumps({"shape": {"B": B, "V": V, "D": D, "G": G, "C_raw": Cr}, "us_stream_launches": t_plain,
                  "us_cuda_graph": t_graph, "graph_replay_bit_identical": same,
                  "MB_per_step": bytes_moved / 1e6, "GBps_graph": bytes_moved / t_graph / 1e3}))
