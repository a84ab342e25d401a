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

Both exchanges are <= 50 KB and latency-bound; on the GPUs of one node they go through ``P2PComm``, the library's
one-kernel all-reduce over NVLink peer memory (``csrc/comm.cu``), with torch.distributed only carrying the 64-byte
IPC handles at start-up.  ``torch.distributed`` collectives remain the checked alternative (CPU tests, multi-node).
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist


class P2PComm:
    """gvcnn_comm (include/gvcnn_b200.h): one-shot all-reduce of small float32 vectors between the GPUs of one
    node - push to every peer's receive buffer over NVLink, flag, wait, add in rank order (bit-identical on every
    rank).  One process per GPU; ``group`` is only used to all-gather the IPC handles."""

    def __init__(self, group=None, device=None):
        from . import _cabi as C
        self._C = C
        self._L = C.lib()
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > C.COMM_MAX_WORLD:
            raise ValueError("P2PComm: at most %d ranks (one node)" % C.COMM_MAX_WORLD)
        have_cuda = torch.cuda.is_available()
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if have_cuda else torch.device("cpu")
        self.device = torch.device(device)
        self._h = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * C.COMM_HANDLE_BYTES)()
        on_gpu = dist.get_backend(group) == "nccl"
        # Construction is COLLECTIVE and fails on every rank or on none: a rank whose buffer / IPC mapping cannot be
        # set up still takes part in both exchanges below, so no rank is left waiting in a collective the others skipped.
        err = None
        import contextlib
        with (torch.cuda.device(self.device) if have_cuda else contextlib.nullcontext()):
            try:
                C.check(self._L.gvcnn_comm_create(ctypes.byref(self._h), self.rank, self.world, handle), "gvcnn_comm_create")
            except Exception as e:                                   # noqa: BLE001
                err = e
            mine = torch.tensor(list(handle) + [0 if err is None else 1], dtype=torch.uint8)
            if on_gpu:
                mine = mine.to(self.device)
            allh = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(allh, mine, group=group)
            table = torch.stack([t.cpu() for t in allh])
            if int(table[:, -1].sum()) != 0 and err is None:
                err = RuntimeError("gvcnn_comm_create failed on another rank")
            if err is None:
                try:
                    C.check(self._L.gvcnn_comm_connect(self._h, bytes(table[:, :-1].reshape(-1).tolist())), "gvcnn_comm_connect")
                except Exception as e:                               # noqa: BLE001
                    err = e
            ok = torch.tensor([0 if err is None else 1], dtype=torch.int32)
            if on_gpu:
                ok = ok.to(self.device)
            dist.all_reduce(ok, group=group)                         # also the barrier: every rank has mapped every peer
            if int(ok.item()) != 0:
                self.close()
                raise RuntimeError("P2PComm could not be set up on every rank: %s"
                                   % (err if err is not None else "peer mapping failed on another rank"))
        # (function, user pointer) in the gvcnn_exchange_fn form the library's entry points take
        self.exchange = (self._L.gvcnn_comm_allreduce_f32, self._h)
        self.exchange_c = (ctypes.cast(self._L.gvcnn_comm_allreduce_f32, ctypes.c_void_p), self._h)

    def all_reduce_(self, t: torch.Tensor, scale: float = 1.0, stream=None):
        """In place: t = (sum over ranks of t) * scale; float32, contiguous, <= COMM_MAX_FLOATS elements,
        ordered on ``stream`` (default: the current stream).  scale = 1 / world is the gradient average."""
        if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
            raise TypeError("P2PComm.all_reduce_: contiguous float32 CUDA tensor expected")
        st = torch.cuda.current_stream(t.device) if stream is None else stream
        self._C.check(self._L.gvcnn_comm_allreduce_scaled_f32(self._h, ctypes.c_void_p(t.data_ptr()), t.numel(),
                                                              ctypes.c_float(scale), ctypes.c_void_p(st.cuda_stream)),
                      "gvcnn_comm_allreduce_scaled_f32")
        return t

    def check(self):
        """Raises if a wait timed out (a peer never arrived).  Synchronises."""
        self._C.check(self._L.gvcnn_comm_error(self._h), "gvcnn_comm_error")

    def close(self):
        if self._h:
            self._L.gvcnn_comm_destroy(self._h)
            self._h = ctypes.c_void_p()


def shard_range(num_shapes: int, rank: int, world_size: int):
    """[lo, hi) of the shapes rank owns: contiguous, sizes differ by at most 1,
    every shape owned exactly once."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size: %d/%d" % (rank, world_size))
    base, rem = divmod(num_shapes, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def steps_per_epoch(num_shapes: int, world_size: int, batch_size: int) -> int:
    """Steps every rank runs per epoch: the step count of the largest shard.  Shard sizes differ by up to one
    shape, so ceil(shard / batch_size) can differ across ranks; ranks must still issue the same number of
    collectives, so the short ones finish the epoch on empty batches."""
    return max(-(-(shard_range(num_shapes, r, world_size)[1] - shard_range(num_shapes, r, world_size)[0]) // batch_size)
               for r in range(world_size))


def global_batch_size(num_shapes: int, world_size: int, batch_size: int, step: int) -> int:
    """Number of shapes all ranks together process at `step` of an epoch (the divisor of the literal batch mean,
    nets/model.py:146, on a sharded batch) - computable locally, no exchange needed."""
    total = 0
    for r in range(world_size):
        lo, hi = shard_range(num_shapes, r, world_size)
        total += max(0, min(batch_size, hi - (lo + step * batch_size)))
    return total


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
