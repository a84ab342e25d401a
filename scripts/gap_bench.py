#!/usr/bin/env python
"""Real-geometry timing (block4 maps [N,10,10,2048], V = 6, num_group = 10): pooling + fusion followed by the
GAP, with the fused map materialised (gvcnn_pool_fuse_fwd/_bwd + a mean) vs folded (gvcnn_pool_fuse_gap_*)."""
import ctypes, json, os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvcnn_tf_b200 import _cabi as C  # noqa: E402
L = C.lib(); dev = torch.device("cuda:0")
N, V, HW, Cc, G = 128, 6, 100, 2048, 10
D = HW * Cc
p = lambda t: ctypes.c_void_p(t.data_ptr()); sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
Fs = [torch.relu(torch.randn(N, V, D, device=dev)) for _ in range(2)]
bins = torch.randint(0, G, (N, V), dtype=torch.int32, device=dev)
status = torch.zeros(4, dtype=torch.int32, device=dev)
S = torch.empty(N, D, device=dev); out = torch.empty(N, Cc, device=dev)
mask = torch.empty(1, N, D, dtype=torch.uint8, device=dev)
dS = torch.randn(N, D, device=dev); dOut = torch.randn(N, Cc, device=dev); dF = torch.empty(N, V, D, device=dev)
wsb = L.gvcnn_pool_fuse_gap_workspace_bytes(N, Cc, HW, C.F32); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
def t(fn, n=30):
    for i in range(3): fn(i)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    for i, (a, c) in enumerate(ev):
        a.record(); fn(i); c.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(c) for a, c in ev) * 1e3
def fwd_plain(i):
    C.check(L.gvcnn_pool_fuse_fwd(p(Fs[i % 2]), p(bins), V, None, 0, p(S), None, p(mask), p(status), N, V, D, G, 0, ctypes.c_float(1.0), 0, 0, sp), "f")
    out.copy_(S.view(N, HW, Cc).mean(dim=1))
def fwd_gap(i):
    C.check(L.gvcnn_pool_fuse_gap_fwd(p(Fs[i % 2]), p(bins), V, p(out), p(mask), p(status), p(ws), wsb, N, V, HW, Cc, G, 0, ctypes.c_float(1.0), 0, 0, sp), "g")
def bwd_plain(i):
    dS.view(N, HW, Cc).copy_((dOut / HW)[:, None, :].expand(N, HW, Cc))
    C.check(L.gvcnn_pool_fuse_bwd(p(dS), p(bins), V, None, 0, p(mask), p(dF), p(status), N, V, D, G, 0, 0, 0, sp), "b")
def bwd_gap(i):
    C.check(L.gvcnn_pool_fuse_gap_bwd(p(dOut), p(bins), V, p(mask), p(dF), p(status), N, V, HW, Cc, G, 0, 0, 0, sp), "gb")
r = {"shape": {"N": N, "V": V, "HW": HW, "C": Cc, "G": G, "F_MB": N * V * D * 4 / 1e6},
     "fwd_materialised_us": t(fwd_plain), "fwd_gap_folded_us": t(fwd_gap),
     "bwd_materialised_us": t(bwd_plain), "bwd_gap_folded_us": t(bwd_gap)}
print(json.dumps(r))
