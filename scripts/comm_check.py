#!/usr/bin/env python
"""torchrun check of the library's one-kernel NVLink all-reduce (csrc/comm.cu, parallel.P2PComm) against
torch.distributed/NCCL: exact equality of the result on every rank for several vector lengths (1 .. 16384 floats,
odd lengths and an unaligned base included), rank-order summation, graph replay, and latency next to NCCL's.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/comm_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvcnn_tf_b200 import parallel  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = parallel.P2PComm()
    out = {"world": world, "cases": []}
    for n in (1, 3, 12, 13, 2047, 2048, 2049, 12300, 16384):
        g = torch.Generator(device=dev).manual_seed(100 * n + rank)
        base = torch.randn(n + 1, generator=g, device=dev)
        for off in (0, 1):                                           # aligned and 4-byte-only aligned base
            x = base[off:off + n].clone() if off == 0 else base[1:1 + n]
            mine = x.clone()
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            want = parts[0].clone()
            for r in range(1, world):
                want = want + parts[r]                               # rank order, one rounding per add
            y = mine.clone() if off == 0 else base[1:1 + n]
            comm.all_reduce_(y if y.is_contiguous() else y.contiguous())
            torch.cuda.synchronize()
            ok = bool(torch.equal(y, want))
            rows = [torch.empty_like(y) for _ in range(world)]
            dist.all_gather(rows, y.contiguous())
            same = all(torch.equal(r_, rows[0]) for r_ in rows)
            out["cases"].append({"n": n, "offset": off, "equals_rank_order_sum": ok, "identical_on_all_ranks": same})
            assert ok and same, (n, off)
    # scaled (gradient average) + many back-to-back calls (double-buffer reuse)
    v = torch.full((12300,), float(rank + 1), device=dev)
    for _ in range(50):
        v.fill_(float(rank + 1))
        comm.all_reduce_(v, 1.0 / world)
    torch.cuda.synchronize()
    assert torch.equal(v, torch.full_like(v, (world + 1) / 2.0)), v[:4]
    # graph capture + replay
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    w = torch.ones(12, device=dev)
    with torch.cuda.stream(side):
        comm.all_reduce_(w, stream=side)
    side.synchronize()
    dist.barrier()
    with torch.cuda.graph(g, stream=side):
        comm.all_reduce_(w, stream=side)
    w.fill_(1.0)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(w, torch.full_like(w, float(world ** 3))), w
    # latency: K back-to-back calls between events
    def lat(fn, k=200):
        for _ in range(10):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / k * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    small, big = torch.zeros(12, device=dev), torch.zeros(12300, device=dev)
    out["us"] = {"p2p_12_floats": lat(lambda: comm.all_reduce_(small)), "nccl_12_floats": lat(lambda: dist.all_reduce(small)),
                 "p2p_12300_floats": lat(lambda: comm.all_reduce_(big)), "nccl_12300_floats": lat(lambda: dist.all_reduce(big))}
    comm.check()
    if rank == 0:
        print(json.dumps(out))
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
