#!/usr/bin/env python
"""BASELINE.json configs[3]: sweep V in {6,12,20,80} x G in {2,4,8,16} x D in {1024,2048} x {fp32,bf16},
B = 4096, both pool modes, per-shape bins.

Every point is timed THE SAME WAY as bench.py's headline: the PDL-chained step (forward = score+bin, pool+fuse;
training = the same with the tie mask + backward) captured into a CUDA graph over rotating input sets (so L2 cannot
serve re-reads) and replayed; per-kernel CUDA-event medians are kept beside it for the table.  bench.py imports
sweep() for the `sweep` key of its N=1 line; run as a script it writes a markdown table + json under gpurun_out/
(copy into profiles/).

    python scripts/sweep.py [--quick] [--out gpurun_out/sweep]
"""
import argparse
import ctypes
import json
import math
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L2_BYTES = 126e6


def _peak():
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        return float(json.load(open(pk))["hbm_gbs"])
    return 6650.0


def sweep(torch, C, L, dev, iters=10, quick=False, per_kernel=False, log=None):
    peak = _peak()
    B, Cr = 4096, 1024
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rows = []
    Vs, Gs, Ds = ([12], [8], [2048]) if quick else ([6, 12, 20, 80], [2, 4, 8, 16], [1024, 2048])
    main = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    sp_side = ctypes.c_void_p(side.cuda_stream)
    sp_main = ctypes.c_void_p(main.cuda_stream)
    for dtype, dt, s in ((torch.float32, C.F32, 4), (torch.bfloat16, C.BF16, 2)):
        for V in Vs:
            for D in Ds:
                fbytes = B * V * D * s
                nsets = max(2, min(4, int(math.ceil(3 * L2_BYTES / fbytes))))
                Fs = [torch.randn(B, V, D, device=dev).to(dtype) for _ in range(nsets)]
                Rs = [torch.randn(B, V, Cr, device=dev).to(dtype) for _ in range(nsets)]
                dSs = [torch.randn(B, D, device=dev).to(dtype) for _ in range(nsets)]
                W = (torch.rand(V, Cr, device=dev) * 2 - 1) * math.sqrt(6.0 / (Cr + 1))
                bias = torch.zeros(V, device=dev)
                scores = torch.empty(B, V, device=dev)
                bins = torch.empty(B, V, dtype=torch.int32, device=dev)
                status = torch.zeros(4, dtype=torch.int32, device=dev)
                Ss = [torch.empty(B, D, device=dev, dtype=dtype) for _ in range(nsets)]
                mask = torch.empty((V + 7) // 8, B, D, dtype=torch.uint8, device=dev)
                dF = torch.empty(B, V, D, device=dev, dtype=dtype)
                for G in Gs:
                    for pool_name, pool in (("max", C.POOL_MAX), ("mean", C.POOL_MEAN)):
                        fill = ctypes.c_float(1.0 if pool_name == "max" else 0.0)
                        want_mask = pool_name == "max"

                        def fwd(i, sp, with_mask):
                            C.check(L.gvcnn_grouping_fusion_fwd(p(Rs[i % nsets]), p(W), p(bias), p(Fs[i % nsets]), None,
                                                                p(scores), p(bins), None, p(Ss[i % nsets]),
                                                                p(mask) if with_mask else None, p(status), B, V, Cr, D, G,
                                                                pool, fill, C.LAYOUT_BVD, C.LAYOUT_BVD, dt, 0, 1, sp), "fwd")

                        def train(i, sp):
                            fwd(i, sp, want_mask)
                            C.check(L.gvcnn_pool_fuse_bwd(p(dSs[i % nsets]), p(bins), V, None, 0, p(mask) if want_mask else None,
                                                          p(dF), p(status), B, V, D, G, pool, C.LAYOUT_BVD, dt, sp), "bwd")

                        def graph_us(step):
                            side.wait_stream(main)
                            with torch.cuda.stream(side):
                                for i in range(nsets):
                                    step(i, sp_side)
                            side.synchronize()
                            g = torch.cuda.CUDAGraph()
                            with torch.cuda.graph(g, stream=side):
                                for i in range(nsets):
                                    step(i, sp_side)
                            torch.cuda.synchronize()
                            g.replay()
                            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            torch.cuda.synchronize()
                            e0.record(main)
                            for _ in range(iters):
                                g.replay()
                            e1.record(main)
                            torch.cuda.synchronize()
                            return e0.elapsed_time(e1) * 1e3 / (iters * nsets)

                        fwd_us = graph_us(lambda i, sp: fwd(i, sp, False))
                        train_us = graph_us(train)
                        a_score = B * (V * Cr * s + 8 * V)
                        a_pool = B * (V * D * s + D * s)
                        row = dict(dtype="bf16" if s == 2 else "fp32", V=V, G=G, D=D, pool=pool_name,
                                   fwd_us=fwd_us, train_us=train_us,
                                   fwd_shapes_per_s=B / (fwd_us * 1e-6), train_shapes_per_s=B / (train_us * 1e-6),
                                   fwd_frac=(a_score + a_pool) / (fwd_us * 1e-6) / 1e9 / peak,
                                   train_frac=(a_score + 2 * a_pool) / (train_us * 1e-6) / 1e9 / peak)
                        if per_kernel:
                            def ev_us(fn, n=iters):
                                for i in range(2):
                                    fn(i)
                                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
                                torch.cuda.synchronize()
                                for i, (a, b_) in enumerate(evs):
                                    a.record()
                                    fn(i)
                                    b_.record()
                                torch.cuda.synchronize()
                                return statistics.median(a.elapsed_time(b_) for a, b_ in evs) * 1e3
                            row["score_us"] = ev_us(lambda i: C.check(L.gvcnn_score_bin_fwd(
                                p(Rs[i % nsets]), p(W), p(bias), None, p(scores), p(bins), None, p(status), B, V, Cr, G,
                                C.LAYOUT_BVD, dt, 0, 1, sp_main), "score"))
                            row["pool_us"] = ev_us(lambda i: C.check(L.gvcnn_pool_fuse_fwd(
                                p(Fs[i % nsets]), p(bins), V, None, 0, p(Ss[i % nsets]), None, None, p(status), B, V, D, G, pool,
                                fill, C.LAYOUT_BVD, dt, sp_main), "pool"))
                            row["pool_mask_us"] = ev_us(lambda i: C.check(L.gvcnn_pool_fuse_fwd(
                                p(Fs[i % nsets]), p(bins), V, None, 0, p(Ss[i % nsets]), None, p(mask), p(status), B, V, D, G,
                                pool, fill, C.LAYOUT_BVD, dt, sp_main), "pool+mask")) if want_mask else row["pool_us"]
                            row["bwd_us"] = ev_us(lambda i: C.check(L.gvcnn_pool_fuse_bwd(
                                p(dSs[i % nsets]), p(bins), V, None, 0, p(mask) if want_mask else None, p(dF), p(status), B, V,
                                D, G, pool, C.LAYOUT_BVD, dt, sp_main), "bwd"))
                        rows.append(row)
                del Fs, Rs, dSs, Ss, mask, dF
                torch.cuda.empty_cache()
            if log:
                log("done %s V=%d" % ("bf16" if s == 2 else "fp32", V))
    if status.tolist()[:2] != [0, 0]:
        raise RuntimeError("sweep: status words report out-of-range / NaN scores")
    return rows


def main():
    import torch
    from gvcnn_tf_b200 import _cabi as C
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default="gpurun_out/sweep")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    peak = _peak()
    L = C.lib()
    dev = torch.device("cuda:0")
    rows = sweep(torch, C, L, dev, iters=args.iters, quick=args.quick, per_kernel=True,
                 log=lambda m: print(m, flush=True))
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump({"B": 4096, "C_raw": 1024, "peak_GBps": peak, "rows": rows}, open(args.out + ".json", "w"))
    fwd = [r["fwd_frac"] for r in rows]
    trn = [r["train_frac"] for r in rows]
    with open(args.out + ".md", "w") as f:
        f.write("# Sweep (BASELINE.json configs[3]): B=4096, C_raw=1024, 1x B200\n\n"
                "fwd / train us = the PDL-chained step replayed from a CUDA graph over rotating inputs (the headline's "
                "method); fractions are algorithmic bytes / that time / measured HBM peak (%.1f GB/s). fwd = score+bin + "
                "pool+fuse; train = the same with the tie mask + backward. Per-kernel columns are single event-bracketed "
                "launches (they include ~4 us of event gap each).\n\n"
                "%d points: fwd frac min %.3f median %.3f; train frac min %.3f median %.3f; %d training points below 0.70.\n\n"
                % (peak, len(rows), min(fwd), statistics.median(fwd), min(trn), statistics.median(trn),
                   sum(1 for t in trn if t < 0.70)))
        f.write("| dtype | V | G | D | pool | fwd us | fwd frac | train us | train frac | score us | pool us | pool+mask us | bwd us |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write("| %s | %d | %d | %d | %s | %.1f | %.3f | %.1f | %.3f | %.1f | %.1f | %.1f | %.1f |\n" % (
                r["dtype"], r["V"], r["G"], r["D"], r["pool"], r["fwd_us"], r["fwd_frac"], r["train_us"], r["train_frac"],
                r["score_us"], r["pool_us"], r["pool_mask_us"], r["bwd_us"]))
    print("wrote", args.out + ".md")


if __name__ == "__main__":
    main()
