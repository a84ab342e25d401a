#!/usr/bin/env python
"""A/B timing of the pooling forward for a few (V, D, dtype, pool) points (B=4096, G=8): CUDA-event time of 60 back-to-back
launches / 60 (best of 3), rotating inputs.  GVCNN_LIB=<path> times an older build of the library on the same box (bisecting regressions)."""
import ctypes, os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvcnn_tf_b200 import _cabi as C  # noqa: E402
if os.environ.get("GVCNN_LIB"):              # an older build of the library, for bisecting
    L = ctypes.CDLL(os.environ["GVCNN_LIB"])
    for name in ("gvcnn_pool_fuse_fwd", "gvcnn_pool_fuse_bwd", "gvcnn_score_bin_fwd"):
        getattr(L, name).restype, getattr(L, name).argtypes = C.SIGNATURES[name]
else:
    L = C.lib()
dev = torch.device("cuda:0"); B, G = 4096, 8
p = lambda t: ctypes.c_void_p(t.data_ptr()); sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
pts = [(6, 1024, "f32", "max"), (6, 2048, "f32", "max"), (12, 2048, "f32", "max"), (20, 1024, "f32", "max"),
       (20, 2048, "f32", "max"), (20, 2048, "f32", "mean"), (6, 2048, "bf16", "max"), (20, 2048, "bf16", "max"),
       (12, 2048, "bf16", "mean")]
for V, D, dt, pool in pts:
    td = torch.float32 if dt == "f32" else torch.bfloat16
    nset = 3
    Fs = [torch.randn(B, V, D, device=dev).to(td) for _ in range(nset)]
    bins = torch.randint(0, G, (B, V), dtype=torch.int32, device=dev)
    S = torch.empty(B, D, device=dev, dtype=td); mask = torch.empty((V + 7) // 8, B, D, dtype=torch.uint8, device=dev)
    status = torch.zeros(4, dtype=torch.int32, device=dev)
    res = []
    for with_mask in (False, True):
        if with_mask and pool == "mean":
            res.append(float("nan")); continue
        def fn(i):
            C.check(L.gvcnn_pool_fuse_fwd(p(Fs[i % nset]), p(bins), V, None, 0, p(S), None, p(mask) if with_mask else None,
                                          p(status), B, V, D, G, 0 if pool == "max" else 1,
                                          ctypes.c_float(1.0 if pool == "max" else 0.0), 0, 0 if dt == "f32" else 1, sp), "p")
        for i in range(5): fn(i)
        # (single-launch event pairs are quantised to ~2 us on this platform: time 60 launches in one bracket)
        best = 1e9
        for _ in range(3):
            a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for i in range(60): fn(i)
            c.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(c) / 60 * 1e3)
        res.append(best)
    # backward (with the tie mask the forward above left behind) and score+bin
    dS = torch.randn(B, D, device=dev).to(td); dF = torch.empty(B, V, D, device=dev, dtype=td)
    Rs = [torch.randn(B, V, 1024, device=dev).to(td) for _ in range(nset)]
    W = (torch.rand(V, 1024, device=dev) * 2 - 1) * 0.0765; bias = torch.zeros(V, device=dev)
    scores = torch.empty(B, V, device=dev); bins2 = torch.empty(B, V, dtype=torch.int32, device=dev)
    def bracket(fn):
        best = 1e9
        for i in range(5): fn(i)
        for _ in range(3):
            a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for i in range(60): fn(i)
            c.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(c) / 60 * 1e3)
        return best
    t_bwd = bracket(lambda i: C.check(L.gvcnn_pool_fuse_bwd(p(dS), p(bins), V, None, 0, p(mask) if pool == "max" else None,
                                                            p(dF), p(status), B, V, D, G, 0 if pool == "max" else 1, 0,
                                                            0 if dt == "f32" else 1, sp), "b"))
    t_score = bracket(lambda i: C.check(L.gvcnn_score_bin_fwd(p(Rs[i % nset]), p(W), p(bias), None, p(scores), p(bins2), None,
                                                              p(status), B, V, 1024, G, 0, 0 if dt == "f32" else 1, 0, 1, sp), "s"))
    print("V=%-3d D=%-5d %-4s %-4s fwd %.1f us  fwd+mask %.1f us  bwd %.1f us  score %.1f us"
          % (V, D, dt, pool, res[0], res[1], t_bwd, t_score), flush=True)
    del Fs, Rs
