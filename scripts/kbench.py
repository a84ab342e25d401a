#!/usr/bin/env python
"""Kernel A/B bench at BASELINE configs[1] size (B=4096, V=12, D=2048, G=8, C_raw=1024, fp32 unless --bf16):
per-kernel back-to-back times (K launches in one event bracket, 3 rotating input AND output sets) and the
graph-replayed forward / training steps.  GVCNN_LIB=<path> loads another build of the library; the GVCNN_* tuning
knobs are read by the library itself.  Prints one JSON line (tag = --tag).

    python scripts/kbench.py --tag base
    GVCNN_BWD_PERSIST=0 python scripts/kbench.py --tag bwd_one_tile
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gvcnn_tf_b200 import _cabi as C  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tag", default="base")
ap.add_argument("--bf16", action="store_true")
ap.add_argument("--V", type=int, default=12)
ap.add_argument("--D", type=int, default=2048)
ap.add_argument("--G", type=int, default=8)
ap.add_argument("--B", type=int, default=4096)
ap.add_argument("--reps", type=int, default=60)
ap.add_argument("--pool", default="max")
args = ap.parse_args()
if os.environ.get("GVCNN_LIB"):
    C.SO_PATH = os.environ["GVCNN_LIB"]
L = C.lib()
dev = torch.device("cuda:0")
B, V, D, G, Cr = args.B, args.V, args.D, args.G, 1024
td, dt, es = (torch.bfloat16, C.BF16, 2) if args.bf16 else (torch.float32, C.F32, 4)
p = lambda t: ctypes.c_void_p(t.data_ptr())
NS = 3
Fs = [torch.randn(B, V, D, device=dev).to(td) for _ in range(NS)]
Rs = [torch.randn(B, V, Cr, device=dev).to(td) for _ in range(NS)]
dSs = [torch.randn(B, D, device=dev).to(td) for _ in range(NS)]
W = (torch.rand(V, Cr, device=dev) * 2 - 1) * 0.0765
bias = torch.zeros(V, device=dev)
bias_lit = (torch.rand(V, device=dev) * 8 - 4)
scores = [torch.empty(B, V, device=dev) for _ in range(NS)]
bins = [torch.empty(B, V, dtype=torch.int32, device=dev) for _ in range(NS)]
xs = [torch.empty(B, V, device=dev) for _ in range(NS)]
xsum = torch.empty(V, device=dev); sc1 = torch.empty(V, device=dev); bins1 = torch.zeros(V, dtype=torch.int32, device=dev)
Ss = [torch.empty(B, D, device=dev, dtype=td) for _ in range(NS)]
masks = [torch.empty((V + 7) // 8, B, D, dtype=torch.uint8, device=dev) for _ in range(NS)]
dFs = [torch.empty(B, V, D, device=dev, dtype=td) for _ in range(NS)]
status = torch.zeros(4, dtype=torch.int32, device=dev)
pool = C.POOL_MAX if args.pool == "max" else C.POOL_MEAN
fill = ctypes.c_float(1.0 if args.pool == "max" else 0.0)
main = torch.cuda.current_stream()
sp = [ctypes.c_void_p(main.cuda_stream)]

K = {
    "score": lambda i: C.check(L.gvcnn_score_bin_fwd(p(Rs[i % NS]), p(W), p(bias), None, p(scores[i % NS]), p(bins[i % NS]), None, p(status), B, V, Cr, G, 0, dt, 0, 1, sp[0]), "score"),
    "score_x": lambda i: C.check(L.gvcnn_view_score_fwd(p(Rs[i % NS]), p(W), p(bias_lit), p(xs[i % NS]), None, B, V, Cr, 0, dt, sp[0]), "score_x"),
    "pool": lambda i: C.check(L.gvcnn_pool_fuse_fwd(p(Fs[i % NS]), p(bins[i % NS]), V, None, 0, p(Ss[i % NS]), None, None, p(status), B, V, D, G, pool, fill, 0, dt, sp[0]), "pool"),
    "pool_mask": lambda i: C.check(L.gvcnn_pool_fuse_fwd(p(Fs[i % NS]), p(bins[i % NS]), V, None, 0, p(Ss[i % NS]), None, p(masks[i % NS]), p(status), B, V, D, G, pool, fill, 0, dt, sp[0]), "pool_mask"),
    "bwd": lambda i: C.check(L.gvcnn_pool_fuse_bwd(p(dSs[i % NS]), p(bins[i % NS]), V, None, 0, p(masks[i % NS]), p(dFs[i % NS]), p(status), B, V, D, G, pool, 0, dt, sp[0]), "bwd"),
    "fwd_shape": lambda i: C.check(L.gvcnn_grouping_fusion_fwd(p(Rs[i % NS]), p(W), p(bias), p(Fs[i % NS]), None, p(scores[i % NS]), p(bins[i % NS]), None, p(Ss[i % NS]), None, p(status), B, V, Cr, D, G, pool, fill, 0, 0, dt, 0, 1, sp[0]), "fwd_shape"),
    "fwd_batch": lambda i: C.check(L.gvcnn_grouping_fusion_batch_fwd(p(Rs[i % NS]), p(W), p(bias_lit), p(Fs[i % NS]), p(xs[i % NS]), p(xsum), None, p(sc1), p(bins1), None, p(Ss[i % NS]), None, p(status), B, V, Cr, D, G, 0, pool, fill, 0, 0, dt, 0, 1, B, None, None, sp[0]), "fwd_batch"),
}


def train_shape(i):
    C.check(L.gvcnn_grouping_fusion_fwd(p(Rs[i % NS]), p(W), p(bias), p(Fs[i % NS]), None, p(scores[i % NS]), p(bins[i % NS]), None, p(Ss[i % NS]), p(masks[i % NS]), p(status), B, V, Cr, D, G, pool, fill, 0, 0, dt, 0, 1, sp[0]), "f")
    K["bwd"](i)


def train_batch(i):
    C.check(L.gvcnn_grouping_fusion_batch_fwd(p(Rs[i % NS]), p(W), p(bias_lit), p(Fs[i % NS]), p(xs[i % NS]), p(xsum), None, p(sc1), p(bins1), None, p(Ss[i % NS]), p(masks[i % NS]), p(status), B, V, Cr, D, G, 0, pool, fill, 0, 0, dt, 0, 1, B, None, None, sp[0]), "f")
    C.check(L.gvcnn_pool_fuse_bwd(p(dSs[i % NS]), p(bins1), 0, None, 0, p(masks[i % NS]), p(dFs[i % NS]), p(status), B, V, D, G, pool, 0, dt, sp[0]), "b")


K["train_shape"], K["train_batch"] = train_shape, train_batch


def b2b(fn, reps):
    for i in range(3):
        fn(i)
    best = 1e9
    for _ in range(3):
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for i in range(reps):
            fn(i)
        c.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(c) / reps * 1e3)
    return best


def graph_us(fn, reps):
    side = torch.cuda.Stream()
    side.wait_stream(main)
    saved = sp[0]
    sp[0] = ctypes.c_void_p(side.cuda_stream)
    with torch.cuda.stream(side):
        for i in range(NS):
            fn(i)
    side.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for i in range(NS):
            fn(i)
    sp[0] = saved
    torch.cuda.synchronize()
    g.replay()
    best = 1e9
    for _ in range(3):
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(reps // NS):
            g.replay()
        c.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(c) / (reps // NS * NS) * 1e3)
    return best


K["score"](0)
out = {"tag": args.tag, "cfg": dict(B=B, V=V, D=D, G=G, dtype="bf16" if args.bf16 else "f32", pool=args.pool)}
out["b2b_us"] = {k: round(b2b(K[k], args.reps), 2) for k in ("score", "score_x", "pool", "pool_mask", "bwd")}
out["graph_us"] = {k: round(graph_us(K[k], args.reps), 2) for k in ("fwd_shape", "fwd_batch", "train_shape", "train_batch")}
assert status.tolist()[:2] == [0, 0]
print(json.dumps(out), flush=True)
