#!/usr/bin/env python
"""BASELINE.json configs[3]: sweep V in {6,12,20,80} x G in {2,4,8,16} x D in {1024,2048} x {fp32,bf16},
B = 4096, both pool modes.  Times each kernel with CUDA events (inputs rotated so L2 cannot serve
re-reads) and writes a markdown table + json under gpurun_out/ (copy into profiles/).

    python scripts/sweep.py [--quick] [--out gpurun_out/sweep]
"""
import argparse
import ctypes
import json
import math
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvcnn_tf_b200 import _cabi as C  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--out", default="gpurun_out/sweep")
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()

peak = 6531.6
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
L = C.lib()
dev = torch.device("cuda:0")
B, Cr = 4096, 1024
p = lambda t: ctypes.c_void_p(t.data_ptr())
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
L2 = 126e6
rows = []
Vs, Gs, Ds = ([12], [8], [2048]) if args.quick else ([6, 12, 20, 80], [2, 4, 8, 16], [1024, 2048])
for dtype, dt, s in ((torch.float32, C.F32, 4), (torch.bfloat16, C.BF16, 2)):
    for V in Vs:
        for D in Ds:
            fbytes = B * V * D * s
            nsets = max(2, int(math.ceil(3 * L2 / fbytes)))
            Fs = [torch.randn(B, V, D, device=dev).to(dtype) for _ in range(nsets)]
            Rs = [torch.randn(B, V, Cr, device=dev).to(dtype) for _ in range(max(2, int(math.ceil(3 * L2 / (B * V * Cr * s)))))]
            dSs = [torch.randn(B, D, device=dev).to(dtype) for _ in range(4)]
            W = (torch.rand(V, Cr, device=dev) * 2 - 1) * math.sqrt(6.0 / (Cr + 1))
            bias = torch.zeros(V, device=dev)
            scores = torch.empty(B, V, device=dev)
            bins = torch.empty(B, V, dtype=torch.int32, device=dev)
            status = torch.zeros(4, dtype=torch.int32, device=dev)
            S = torch.empty(B, D, device=dev, dtype=dtype)
            mask = torch.empty((V + 7) // 8, B, D, dtype=torch.uint8, device=dev)
            dF = torch.empty(B, V, D, device=dev, dtype=dtype)
            for G in Gs:
                def timeit(fn, n=args.iters):
                    for i in range(3):
                        fn(i)
                    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
                    torch.cuda.synchronize()
                    for i, (a, b_) in enumerate(evs):
                        a.record()
                        fn(i)
                        b_.record()
                    torch.cuda.synchronize()
                    return statistics.median(a.elapsed_time(b_) for a, b_ in evs) * 1e3   # us

                t_score = timeit(lambda i: C.check(L.gvcnn_score_bin_fwd(p(Rs[i % len(Rs)]), p(W), p(bias), None, p(scores), p(bins), None,
                                                                        p(status), B, V, Cr, G, C.LAYOUT_BVD, dt, 0, 1, sp), "score"))
                for pool_name, pool in (("max", C.POOL_MAX), ("mean", C.POOL_MEAN)):
                    fill = ctypes.c_float(1.0 if pool_name == "max" else 0.0)
                    t_fwd = timeit(lambda i: C.check(L.gvcnn_pool_fuse_fwd(p(Fs[i % nsets]), p(bins), V, None, 0, p(S), None, None, p(status),
                                                                          B, V, D, G, pool, fill, C.LAYOUT_BVD, dt, sp), "fwd"))
                    t_fwdm = timeit(lambda i: C.check(L.gvcnn_pool_fuse_fwd(p(Fs[i % nsets]), p(bins), V, None, 0, p(S), None, p(mask), p(status),
                                                                           B, V, D, G, pool, fill, C.LAYOUT_BVD, dt, sp), "fwd+mask")) \
                        if pool_name == "max" else t_fwd
                    t_bwd = timeit(lambda i: C.check(L.gvcnn_pool_fuse_bwd(p(dSs[i % 4]), p(bins), V, None, 0, p(mask), p(dF), p(status),
                                                                          B, V, D, G, pool, C.LAYOUT_BVD, dt, sp), "bwd"))
                    a_score = B * (V * Cr * s + 8 * V)
                    a_pool = B * (V * D * s + D * s)
                    fwd_us = t_score + t_fwd
                    train_us = t_score + t_fwdm + t_bwd
                    rows.append(dict(dtype="bf16" if s == 2 else "fp32", V=V, G=G, D=D, pool=pool_name,
                                     score_us=t_score, fwd_us=t_fwd, fwd_mask_us=t_fwdm, bwd_us=t_bwd,
                                     fwd_shapes_per_s=B / (fwd_us * 1e-6), train_shapes_per_s=B / (train_us * 1e-6),
                                     fwd_frac=(a_score + a_pool) / (fwd_us * 1e-6) / 1e9 / peak,
                                     train_frac=(a_score + 2 * a_pool) / (train_us * 1e-6) / 1e9 / peak,
                                     pool_fwd_frac=a_pool / (t_fwd * 1e-6) / 1e9 / peak,
                                     bwd_frac=a_pool / (t_bwd * 1e-6) / 1e9 / peak,
                                     score_frac=a_score / (t_score * 1e-6) / 1e9 / peak))
            del Fs, Rs, dSs, S, mask, dF
            torch.cuda.empty_cache()
        print("done", "bf16" if s == 2 else "fp32", "V", V, flush=True)
assert status.tolist()[:2] == [0, 0]
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
json.dump({"B": B, "C_raw": Cr, "peak_GBps": peak, "rows": rows}, open(args.out + ".json", "w"))
with open(args.out + ".md", "w") as f:
    f.write("# Sweep (BASELINE.json configs[3]): B=4096, C_raw=1024, per-kernel CUDA-event medians, 1x B200\n\n"
            "Fractions are algorithmic bytes / time / measured HBM peak (%.1f GB/s). fwd = score+bin + pool+fuse; "
            "train = score+bin + pool+fuse with tie mask + backward.\n\n" % peak)
    f.write("| dtype | V | G | D | pool | score us | fwd us | fwd+mask us | bwd us | fwd Mshapes/s | fwd frac | train Mshapes/s | train frac |\n")
    f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for r in rows:
        f.write("| %s | %d | %d | %d | %s | %.1f | %.1f | %.1f | %.1f | %.2f | %.3f | %.2f | %.3f |\n" % (
            r["dtype"], r["V"], r["G"], r["D"], r["pool"], r["score_us"], r["fwd_us"], r["fwd_mask_us"], r["bwd_us"],
            r["fwd_shapes_per_s"] / 1e6, r["fwd_frac"], r["train_shapes_per_s"] / 1e6, r["train_frac"]))
print("wrote", args.out + ".md")
