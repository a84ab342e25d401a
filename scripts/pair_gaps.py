#!/usr/bin/env python
"""Where does the training step lose time between kernels?  Times sequences of the three kernels of the step
(S = score+bin, P = pool+fuse with tie mask, p = pool+fuse without, B = backward, q/Q = p/P through the one-tile-per-CTA forward kernel) back to back, 60 repetitions per
event bracket, rotating inputs, at BASELINE configs[1] size, and prints the time per repetition next to the sum of
the single-kernel back-to-back times."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvcnn_tf_b200 import _cabi as C  # noqa: E402
L = C.lib(); dev = torch.device("cuda:0"); B, V, D, G, Cr = 4096, 12, 2048, 8, 1024
p = lambda t: ctypes.c_void_p(t.data_ptr()); sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
nset = 3
Fs = [torch.randn(B, V, D, device=dev) for _ in range(nset)]
Rs = [torch.randn(B, V, Cr, device=dev) for _ in range(nset)]
dSs = [torch.randn(B, D, device=dev) for _ in range(nset)]
W = (torch.rand(V, Cr, device=dev) * 2 - 1) * 0.0765; bias = torch.zeros(V, device=dev)
scores = torch.empty(B, V, device=dev); bins = torch.empty(B, V, dtype=torch.int32, device=dev)
S = torch.empty(B, D, device=dev); mask = torch.empty((V + 7) // 8, B, D, dtype=torch.uint8, device=dev)
dF = torch.empty(B, V, D, device=dev); status = torch.zeros(4, dtype=torch.int32, device=dev)
one = ctypes.c_float(1.0)
K = {
    "S": lambda i: C.check(L.gvcnn_score_bin_fwd(p(Rs[i % nset]), p(W), p(bias), None, p(scores), p(bins), None, p(status), B, V, Cr, G, 0, 0, 0, 1, sp), "s"),
    "P": lambda i: C.check(L.gvcnn_pool_fuse_fwd(p(Fs[i % nset]), p(bins), V, None, 0, p(S), None, p(mask), p(status), B, V, D, G, 0, one, 0, 0, sp), "P"),
    "p": lambda i: C.check(L.gvcnn_pool_fuse_fwd(p(Fs[i % nset]), p(bins), V, None, 0, p(S), None, None, p(status), B, V, D, G, 0, one, 0, 0, sp), "p"),
    "B": lambda i: C.check(L.gvcnn_pool_fuse_bwd(p(dSs[i % nset]), p(bins), V, None, 0, p(mask), p(dF), p(status), B, V, D, G, 0, 0, 0, sp), "b"),
}
def generic_pool(i, with_mask):   # the one-tile-per-CTA kernel (variant 1) instead of the persistent ring
    C.check(L.gvcnn_pool_fuse_fwd(p(Fs[i % nset]), p(bins), V, None, 0, p(S), None, p(mask) if with_mask else None, p(status), B, V, D, G, C.pool_variant(0, 1), one, 0, 0, sp), "q")
K["q"] = lambda i: generic_pool(i, False)
K["Q"] = lambda i: generic_pool(i, True)
K["S"](0); torch.cuda.synchronize()
def seq_time(seq, reps=60):
    best = 1e9
    for i in range(3):
        for k in seq: K[k](i)
    for _ in range(3):
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for i in range(reps):
            for k in seq: K[k](i)
        c.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(c) / reps * 1e3)
    return best
single = {k: seq_time(k) for k in "SPpBqQ"}
print("single-kernel back to back:", {k: round(v, 1) for k, v in single.items()})
for seq in ("Sp", "SP", "PB", "BS", "SPB", "pB", "Sq", "qB", "SpB", "BpS", "ppB", "pBB", "ppBB"):
    t = seq_time(seq)
    print("%-4s %.1f us   sum of singles %.1f   extra %.1f" % (seq, t, sum(single[k] for k in seq), t - sum(single[k] for k in seq)))
