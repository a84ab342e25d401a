#!/usr/bin/env python
"""How much of each kernel's time is fixed (launch ramp + tail) vs proportional to the batch:
times score+bin and pool+fuse forward at B = 1024 ... 16384 (V=12, D=2048, G=8, fp32) and fits t = a + b*B."""
import ctypes, json, os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvcnn_tf_b200 import _cabi as C  # noqa: E402
L = C.lib(); dev = torch.device("cuda:0"); V, D, G, Cr = 12, 2048, 8, 1024
p = lambda t: ctypes.c_void_p(t.data_ptr()); sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
out = {}
for B in (1024, 2048, 4096, 8192, 16384):
    nset = max(2, (400 << 20) // (B * V * D * 4) + 1)
    Fs = [torch.randn(B, V, D, device=dev) for _ in range(nset)]
    Rs = [torch.randn(B, V, Cr, device=dev) for _ in range(nset)]
    W = (torch.rand(V, Cr, device=dev) * 2 - 1) * 0.0765; b = torch.zeros(V, device=dev)
    scores = torch.empty(B, V, device=dev); bins = torch.empty(B, V, dtype=torch.int32, device=dev)
    status = torch.zeros(4, dtype=torch.int32, device=dev); S = torch.empty(B, D, device=dev)
    def t(fn, n=30):
        for i in range(3): fn(i)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        torch.cuda.synchronize()
        for i, (a, c) in enumerate(ev):
            a.record(); fn(i); c.record()
        torch.cuda.synchronize()
        return statistics.median(a.elapsed_time(c) for a, c in ev) * 1e3
    ts = t(lambda i: C.check(L.gvcnn_score_bin_fwd(p(Rs[i % nset]), p(W), p(b), None, p(scores), p(bins), None, p(status), B, V, Cr, G, 0, 0, 0, 1, sp), "s"))
    tp = t(lambda i: C.check(L.gvcnn_pool_fuse_fwd(p(Fs[i % nset]), p(bins), V, None, 0, p(S), None, None, p(status), B, V, D, G, 0, ctypes.c_float(1.0), 0, 0, sp), "p"))
    out[B] = (ts, tp)
    print(B, "score %.1f us  pool %.1f us" % (ts, tp), flush=True)
    del Fs, Rs
Bs = sorted(out)
for name, idx in (("score", 0), ("pool", 1)):
    x0, x1 = Bs[0], Bs[-1]
    slope = (out[x1][idx] - out[x0][idx]) / (x1 - x0)
    a = out[x0][idx] - slope * x0
    byt = (V * Cr * 4 + 8 * V) if idx == 0 else (V * D * 4 + D * 4)
    print("%s: fixed %.1f us, marginal %.4f us/shape = %.0f GB/s" % (name, a, slope, byt / slope / 1e3))
