#!/usr/bin/env python
"""gvcnn_gap_score_bin_fwd (GAP of the block3 maps folded into the score kernel, nets/model.py:144-145) at the
reference's geometry: N shapes x V views of 10 x 10 x 1024 maps.  Back-to-back launches over rotating inputs."""
import argparse, ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvcnn_tf_b200 import _cabi as C  # noqa: E402
ap = argparse.ArgumentParser(); ap.add_argument("--N", type=int, default=128); ap.add_argument("--V", type=int, default=6)
ap.add_argument("--bf16", action="store_true"); args = ap.parse_args()
if os.environ.get("GVCNN_LIB"):
    C.SO_PATH = os.environ["GVCNN_LIB"]
L = C.lib(); dev = torch.device("cuda:0"); N, V, HW, Cr, G = args.N, args.V, 100, 1024, 10
td, dt, es = (torch.bfloat16, C.BF16, 2) if args.bf16 else (torch.float32, C.F32, 4)
p = lambda t: ctypes.c_void_p(t.data_ptr()); sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
NS = max(2, int(3 * 126e6 / (N * V * HW * Cr * es)) + 1)
maps = [torch.randn(N, V, HW, Cr, device=dev).to(td) for _ in range(NS)]
W = (torch.rand(V, Cr, device=dev) * 2 - 1) * 0.0765; bias = torch.zeros(V, device=dev)
sc = torch.empty(N, V, device=dev); bi = torch.empty(N, V, dtype=torch.int32, device=dev); st = torch.zeros(4, dtype=torch.int32, device=dev)
fn = lambda i: C.check(L.gvcnn_gap_score_bin_fwd(p(maps[i % NS]), p(W), p(bias), None, None, p(sc), p(bi), None, p(st), N, V, HW, Cr, G, 0, dt, 1, 0, 1, sp), "gap")
for i in range(5): fn(i)
best = 1e9
for _ in range(3):
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for i in range(60): fn(i)
    c.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(c) / 60 * 1e3)
# the eager alternative the kernel replaces: torch mean over positions, then the score kernel
R = torch.empty(N, V, Cr, device=dev, dtype=td)
def eager(i):
    torch.mean(maps[i % NS].float(), dim=2, out=None).to(td)
for i in range(3): eager(i)
a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for i in range(20): eager(i)
c.record(); torch.cuda.synchronize(); t_eager = a.elapsed_time(c) / 20 * 1e3
nbytes = N * V * HW * Cr * es
peak = 6531.6
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk): peak = float(json.load(open(pk))["hbm_gbs"])
print(json.dumps({"N": N, "V": V, "HW": HW, "C_raw": Cr, "dtype": "bf16" if args.bf16 else "f32", "bytes": nbytes, "us": best,
                  "GBps": nbytes / best / 1e3, "frac_of_peak": nbytes / best / 1e3 / peak, "eager_mean_only_us": t_eager}))
