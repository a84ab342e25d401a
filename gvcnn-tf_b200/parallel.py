"""Multi-GPU plumbing for the path: one process per GPU, shapes sharded across ranks.

The path is embarrassingly parallel over shapes: rank r of K
owns a contiguous slice of the batch and runs the same two kernels on it; the
forward needs no exchange in per-shape mode.  The only collectives are
  * training: one all-reduce (sum, then / K) of the flat parameter-gradient
    bucket - the B200 form of ``nccl_ops.all_sum`` + ``* 1/K`` in the
    reference's (dead) ``utils/_train_helper.py:17-31``;
  * literal ``score_reduce='batch'``: a pre-binning all-reduce of V partial
    sums (``model.score_bin(process_group=...)``);
  * init: broadcast of the head's parameters from rank 0
    (``utils/_train_helper.py:66-94``).
Works with the ``nccl`` backend on GPUs and ``gloo`` on CPU (tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(num_shapes: int, rank: int, world_size: int):
    """[lo, hi) of the shapes rank owns: contiguous, sizes differ by at most 1,
    every shape owned exactly once."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size: %d/%d" % (rank, world_size))
    base, rem = divmod(num_shapes, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None):
    """Tower-0 -> all copy of the variables at init (utils/_train_helper.py:66-94)."""
    for p in module.parameters():
        dist.broadcast(p.data, src=src, group=group)
    for b in module.buffers():
        dist.broadcast(b.data, src=src, group=group)


class GradBucket:
    """One flat float32 bucket for the head's parameter gradients, all-reduced
    with a single collective (payload V*(C_raw+1) floats = 49 KB at V=12 plus the
    classifier; latency-bound, so one launch instead of one per variable)."""

    def __init__(self, params, device=None):
        self.params = [p for p in params]
        n = sum(p.numel() for p in self.params)
        dev = device if device is not None else (self.params[0].device if self.params else "cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)

    def pack(self):
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:        # FC-score params get no gradient in literal mode
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        return self.flat

    def all_reduce_mean(self, group=None, async_op=False):
        """sum over ranks then * 1/K; returns the work handle when async_op."""
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        self._scale = 1.0 / dist.get_world_size(group)
        if async_op:
            return work
        self.flat.mul_(self._scale)
        return None

    def finish(self, work=None):
        if work is not None:
            work.wait()
            self.flat.mul_(self._scale)

    def unpack(self):
        off = 0
        for p in self.params:
            n = p.numel()
            g = self.flat[off:off + n].reshape(p.shape)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
