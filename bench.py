#!/usr/bin/env python
"""bench.py - GVCNN grouping + fusion shapes/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]    # the reference's op graph on host cores

Workload (config.workload): BASELINE.json configs[1] - grouping + fusion FORWARD, 12 views,
D = 2048, G = 8 groups, B = 4096 synthetic shapes per GPU, C_raw = 1024, fp32, per-shape scores
(SURVEY.md 8d).  One step = score+bin kernel, then pool+fuse kernel, over one batch.  N > 1: one
process per GPU (torchrun), shapes sharded by rank, no data-path collective ("weak" scaling: 4096
shapes per GPU).  The same run also measures the training step (configs[2]: forward with tie mask +
backward + the parameter-gradient all-reduce) and reports it under "fwd_bwd".

One JSON line on stdout (rank 0).  `value` = device-resident whole-job shapes/s; `e2e` = the same
through the host-buffer C-ABI entry point (pinned host buffers, H2D + kernels + D2H every step);
`roofline` = the pool+fuse kernel's algorithmic bytes / its CUDA-event duration vs the measured HBM
peak; `cpu_baseline` = the reference's op sequence restated in torch-CPU, timed on this box's cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "grouping+fusion shapes/s (12-view, D=2048)"
UNIT = "shapes/s"
CFG = dict(B=4096, V=12, D=2048, G=8, C_raw=1024, pool="max", empty_fill=1.0, score_reduce="shape")


# --------------------------------------------------------------------------- helpers
def algorithmic_bytes(B, V, D, C, s):
    """SURVEY.md 8d / BASELINE.md 3: every compulsory tensor counted once."""
    score = B * (V * C * s + 8 * V)
    pool = B * (V * D * s + D * s)
    bwd = B * (D * s + V * D * s)
    return {"score": score, "pool_fwd": pool, "fwd": score + pool, "bwd": bwd, "fwd_bwd": score + pool + bwd}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed regions run."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:                                           # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:                                   # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:                                       # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML reports as local to its GPU, so the pinned host buffers of the
    end-to-end leg are first-touched on the NUMA node whose PCIe root the GPU hangs off (matters when 8
    ranks stream 50 GB/s each from host memory).  Best effort: silently skipped if NVML says nothing."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:                                               # noqa: BLE001
        pass


def synth_inputs(B, V, D, C, seed_base, rank):
    """CPU generators with fixed seeds so oracle and GPU see identical bits."""
    import torch
    off = 1000 * rank
    F = torch.randn((B, V, D), generator=torch.Generator().manual_seed(seed_base + 0 + off))
    R = torch.randn((B, V, C), generator=torch.Generator().manual_seed(seed_base + 1 + off))
    lim = math.sqrt(6.0 / (C + 1))                                   # Keras glorot-uniform of Dense(1)
    W = (torch.rand((V, C), generator=torch.Generator().manual_seed(2)) * 2 - 1) * lim
    b = torch.zeros(V)
    dS = torch.randn((B, D), generator=torch.Generator().manual_seed(seed_base + 3 + off))
    return F, R, W, b, dS


# --------------------------------------------------------------------------- CPU baseline / reference arm
def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_reference_time(B, steps, warmup):
    """The reference's own op sequence for this path on host cores: oracle/gvcnn_oracle_torch
    .reference_step_cpu (stack views -> per group: where -> gather | ones dummy -> reduce_max ->
    multiply -> add_n -> div, after scores -> host binning -> weights; train.py:270-288,
    nets/model.py:16-102).  Returns (seconds per step list, threads)."""
    import torch
    from oracle import gvcnn_oracle_torch as OT
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    V, D, C, G = CFG["V"], CFG["D"], CFG["C_raw"], CFG["G"]
    F, R, W, b, _ = synth_inputs(B, V, D, C, 0, 0)
    b = (torch.rand(V, generator=torch.Generator().manual_seed(9)) * 8 - 4)   # spread the batch-mean bins
    views = [F[:, v, :].contiguous() for v in range(V)]              # the reference's list of V view tensors
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        S = OT.reference_step_cpu(views, R, W, b, G, pool=CFG["pool"], empty_fill=CFG["empty_fill"])
        float(S[0, 0])
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, threads


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    B = 512                                                          # bounded sample of the B=4096 batch per step
    times, threads = cpu_reference_time(B, args.steps, args.warmup)
    total = sum(times)
    value = B * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "grouping+fusion forward, 12 views, D=2048, G=8, C_raw=1024, fp32 "
                               "(BASELINE.json configs[1]); each step = a bounded sample of %d of the 4096 shapes" % B,
                   "B_per_step": B, "V": CFG["V"], "D": CFG["D"], "G": CFG["G"], "C_raw": CFG["C_raw"],
                   "pool": CFG["pool"], "empty_fill": CFG["empty_fill"], "score_reduce": "batch (reference-literal)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "cpu_model": cpu_model(),
                         "sample": "%d steps x %d shapes; reference op graph restated in torch-CPU "
                                   "(TensorFlow 1.x is not installable in this image)" % (len(times), B)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- CUDA arm
def run_cuda_arm(args):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device. The product path is sm_100a CUDA only (no CPU fallback); "
                         "use --impl reference for the host baseline.")
    from gvcnn_tf_b200 import _cabi as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the gradient all-reduce must get SM slots while the dF kernel is running: high-priority NCCL stream
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    torch.cuda.set_device(dev)
    bind_to_gpu_numa_node(dev.index if dev.index is not None else 0)
    L = C.lib()
    C.check(L.gvcnn_check_device(), "gvcnn_check_device")

    B, V, D, G, Cr = CFG["B"], CFG["V"], CFG["D"], CFG["G"], CFG["C_raw"]
    if args.scaling == "strong":                                     # B_total fixed at 4096
        if B % world:
            raise SystemExit("bench.py: --scaling strong needs %d %% n_gpus == 0" % B)
        B //= world
    K, Wm = args.steps, args.warmup
    s = 4
    NSETS = 2 if args.scaling == "weak" else 2 * world               # rotating input sets: 1.2 GB in total >> 126 MB L2
    sets = []
    host = None
    for i in range(NSETS):
        F, R, Wt, bt, dS = synth_inputs(B, V, D, Cr, 10 * i, rank)
        if i == 0:
            host = (F, R, dS)
        sets.append((F.to(dev), R.to(dev), dS.to(dev)))
    Wd, bd = Wt.to(dev), bt.to(dev)
    scores = torch.empty((B, V), dtype=torch.float32, device=dev)
    bins = torch.empty((B, V), dtype=torch.int32, device=dev)
    status = torch.zeros(C.STATUS_WORDS, dtype=torch.int32, device=dev)
    S = torch.empty((B, D), dtype=torch.float32, device=dev)
    mask = torch.empty(((V + 7) // 8, B, D), dtype=torch.uint8, device=dev)
    dF = torch.empty((B, V, D), dtype=torch.float32, device=dev)
    grad_bucket = torch.zeros(V * (Cr + 1), dtype=torch.float32, device=dev)   # FC-score grads (zeros: SURVEY D6)
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    pool = C.POOL_MAX if CFG["pool"] == "max" else C.POOL_MEAN
    fill = ctypes.c_float(CFG["empty_fill"])

    def k_score(Rd):
        C.check(L.gvcnn_score_bin_fwd(p(Rd), p(Wd), p(bd), None, p(scores), p(bins), None, p(status),
                                      B, V, Cr, G, C.LAYOUT_BVD, C.F32, 0, 1, sp), "score_bin_fwd")

    def k_pool(Fd, with_mask):
        C.check(L.gvcnn_pool_fuse_fwd(p(Fd), p(bins), V, None, 0, p(S), None, p(mask) if with_mask else None,
                                      p(status), B, V, D, G, pool, fill, C.LAYOUT_BVD, C.F32, sp), "pool_fuse_fwd")

    def k_bwd(dSd):
        C.check(L.gvcnn_pool_fuse_bwd(p(dSd), p(bins), V, None, 0, p(mask), p(dF), p(status),
                                      B, V, D, G, pool, C.LAYOUT_BVD, C.F32, sp), "pool_fuse_bwd")

    def k_fwd(Rd, Fd, with_mask):
        # the product's forward entry point: score+bin then pool+fuse, chained with programmatic dependent launch
        C.check(L.gvcnn_grouping_fusion_fwd(p(Rd), p(Wd), p(bd), p(Fd), None, p(scores), p(bins), None, p(S),
                                            p(mask) if with_mask else None, p(status), B, V, Cr, D, G, pool, fill,
                                            C.LAYOUT_BVD, C.LAYOUT_BVD, C.F32, 0, 1, sp), "grouping_fusion_fwd")

    fwd_launches = 2

    def step_fwd(i):
        Fd, Rd, _ = sets[i % NSETS]
        k_fwd(Rd, Fd, False)

    def step_train(i):
        Fd, Rd, dSd = sets[i % NSETS]
        k_fwd(Rd, Fd, True)
        # sum then * 1/K in one collective (ReduceOp.AVG); launched before, and overlapping, the dF kernel
        work = dist.all_reduce(grad_bucket, op=dist.ReduceOp.AVG, async_op=True) if world > 1 else None
        k_bwd(dSd)
        if work is not None:
            work.wait()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_graph(step_fn):
        """Captures NSETS consecutive steps (one per rotating input set) into a CUDA graph so the timed loop
        does not depend on how fast this box's CPU can issue launches.  Returns None if capture is refused."""
        try:
            side = torch.cuda.Stream()
            side.wait_stream(stream)
            g = torch.cuda.CUDAGraph()
            sp_side = ctypes.c_void_p(side.cuda_stream)
            nonlocal sp
            saved = sp
            sp = sp_side
            try:
                with torch.cuda.stream(side):
                    for i in range(NSETS):
                        step_fn(i)                                   # warm the capture stream
                side.synchronize()
                with torch.cuda.graph(g, stream=side):
                    for i in range(NSETS):
                        step_fn(i)
            finally:
                sp = saved
            torch.cuda.synchronize()
            return g
        except Exception:                                            # noqa: BLE001
            torch.cuda.synchronize()
            return None

    def timed(step_fn, k, graph=None):
        """K steps bracketed by barrier + synchronize; device time by CUDA events; max over ranks.
        With a graph: k // NSETS replays of the NSETS-step graph plus k % NSETS eager steps = exactly k steps."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if graph is not None:
            for _ in range(k // NSETS):
                graph.replay()
            for i in range(k % NSETS):
                step_fn(i)
        else:
            for i in range(k):
                step_fn(i)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(dev.index if dev.index is not None else 0) if rank == 0 else None

    # ---- warm-up, then the headline region: exactly K forward steps
    for i in range(max(Wm, 3)):
        step_fwd(i)
        step_train(i)
    graph_fwd = None if args.no_graph else make_graph(step_fwd)
    graph_train = None if (args.no_graph or world > 1) else make_graph(step_train)   # NCCL stays outside graphs
    if sampler:
        sampler.start()
    ms_fwd = timed(step_fwd, K, graph_fwd)

    # ---- per-kernel durations, measured live with events around each launch (second region so the
    #      events do not sit inside the headline number)
    barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    for i in range(K):
        Fd, Rd, _ = sets[i % NSETS]
        evs[i][0].record(stream)
        k_score(Rd)
        evs[i][1].record(stream)
        k_pool(Fd, False)
        evs[i][2].record(stream)
    barrier()
    t_score = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
    t_pool = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)
    # ... the pool kernel alone, K launches back to back in one event bracket (no per-launch event gaps)
    ms_pool_b2b = timed(lambda i: k_pool(sets[i % NSETS][0], False), K)
    # ... and of the fused forward launch used by the headline region
    barrier()
    evf = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    for i in range(K):
        Fd, Rd, _ = sets[i % NSETS]
        evf[i][0].record(stream)
        k_fwd(Rd, Fd, False)
        evf[i][1].record(stream)
    barrier()
    t_fused = statistics.mean(e[0].elapsed_time(e[1]) for e in evf)

    # ---- training step (configs[2]): fwd with tie mask + bwd (+ grad all-reduce when N > 1)
    ms_train = timed(step_train, K, graph_train)
    barrier()
    evb = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    for i in range(K):
        Fd, Rd, dSd = sets[i % NSETS]
        evb[i][0].record(stream)
        k_score(Rd)
        evb[i][1].record(stream)
        k_pool(Fd, True)
        evb[i][2].record(stream)
        k_bwd(dSd)
        evb[i][3].record(stream)
    barrier()
    t_pool_m = statistics.mean(e[1].elapsed_time(e[2]) for e in evb)
    t_bwd = statistics.mean(e[2].elapsed_time(e[3]) for e in evb)
    # ... and the gradient all-reduce alone (N > 1): K back-to-back collectives, nothing to hide behind
    us_allreduce = None
    if world > 1:
        def step_allreduce(i):
            dist.all_reduce(grad_bucket, op=dist.ReduceOp.AVG)
        for i in range(3):
            step_allreduce(i)
        us_allreduce = timed(step_allreduce, K) / K * 1e3

    # ---- the reference-literal mode (one scheme per batch, nets/model.py:146): x per (shape, view), deterministic
    #      column sums, V scores/bins, pooling with the shared bin row.  Reported beside the per-shape headline.
    xb = torch.empty((B, V), dtype=torch.float32, device=dev)
    xsum = torch.empty((1, V), dtype=torch.float32, device=dev)
    sc1 = torch.empty((1, V), dtype=torch.float32, device=dev)
    bins1 = torch.empty((1, V), dtype=torch.int32, device=dev)
    bias_lit = (torch.rand(V, generator=torch.Generator().manual_seed(9)) * 8 - 4).to(dev)   # spread the batch means

    def step_literal(i):
        Fd, Rd, _ = sets[i % NSETS]
        C.check(L.gvcnn_view_score_fwd(p(Rd), p(Wd), p(bias_lit), p(xb), B, V, Cr, C.LAYOUT_BVD, C.F32, sp), "view_score_fwd")
        C.check(L.gvcnn_batch_sum_x(p(xb), p(xsum), B, V, sp), "batch_sum_x")
        C.check(L.gvcnn_score_bin(p(xsum), ctypes.c_float(float(B)), p(sc1), p(bins1), None, p(status), V, G, 0, 1, sp),
                "score_bin")
        C.check(L.gvcnn_pool_fuse_fwd(p(Fd), p(bins1), 0, None, 0, p(S), None, None, p(status), B, V, D, G, pool, fill,
                                      C.LAYOUT_BVD, C.F32, sp), "pool_fuse_fwd")

    for i in range(3):
        step_literal(i)
    ms_literal = timed(step_literal, K)

    # ---- end to end through the host-buffer C-ABI entry point (pinned host memory)
    Ke = max(1, min(K, args.e2e_steps))
    Fh, Rh, dSh = (t.pin_memory() for t in host)
    Sh = torch.empty((B, D), dtype=torch.float32).pin_memory()
    bins_h = torch.empty((B, V), dtype=torch.int32).pin_memory()
    st_h = torch.zeros(C.STATUS_WORDS, dtype=torch.int32)
    chunk = args.e2e_chunk
    ws_bytes = L.gvcnn_host_workspace_bytes(chunk, V, Cr, D, C.F32, 0)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)

    def e2e_step():
        C.check(L.gvcnn_grouping_fusion_host(p(Rh), p(Fh), p(Wd), p(bd), p(Sh), None, p(bins_h), None, None,
                                             p(st_h), B, V, Cr, D, G, pool, fill, C.F32, chunk, p(ws), ws_bytes),
                "gvcnn_grouping_fusion_host")

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()                                                   # synchronous: returns with S in host memory
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop() if sampler else None
    h2d = B * V * (Cr + D) * s
    d2h = B * D * s + B * V * 4

    # ---- sanity: the timed path produced the oracle's bins-consistent result on this rank
    st = status.tolist()
    if any(st[:2]):
        raise SystemExit("bench.py: status words report out-of-range/NaN scores: %s" % st)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    ab = algorithmic_bytes(B, V, D, Cr, s)
    peak, peak_src = measured_peaks()
    value = world * B * K / (ms_fwd * 1e-3)
    ach_pool = ab["pool_fwd"] / (t_pool * 1e-3) / 1e9
    ach_fused = ab["fwd"] / (t_fused * 1e-3) / 1e9
    ach_score = ab["score"] / (t_score * 1e-3) / 1e9
    ach_bwd = ab["bwd"] / (t_bwd * 1e-3) / 1e9
    ach_fwd_step = ab["fwd"] / (ms_fwd / K * 1e-3) / 1e9
    ach_train_step = ab["fwd_bwd"] / (ms_train / K * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")               # ncu dram bytes of the pool kernel, if captured
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                traffic = json.load(f).get("fused_fwd_dram_bytes_per_launch" if fwd_launches == 1
                                           else "pool_fuse_fwd_dram_bytes_per_launch")
        except Exception:                                            # noqa: BLE001
            traffic = None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(Wm, 3),
        "ms_per_step": ms_fwd / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "grouping+fusion forward (score+bin, pool+fuse), 12 views, D=2048, G=8, "
                               "B=%d shapes per GPU, C_raw=1024, fp32 (BASELINE.json configs[1]%s)"
                               % (B, "" if args.scaling == "weak" else "; strong scaling: 4096 shapes split over the ranks"),
                   "B_per_gpu": B, "V": V, "D": D, "G": G, "C_raw": Cr, "pool": CFG["pool"],
                   "empty_fill": CFG["empty_fill"], "score_reduce": "shape", "parallelism": "shape-sharded x%d" % world,
                   "launch": ("CUDA graph of %d steps replayed K/%d times" % (NSETS, NSETS)) if graph_fwd is not None
                             else "stream launches",
                   "l2": "inputs larger than L2 (F %.0f MB + R %.0f MB per step vs 126 MB) and %d rotating input sets"
                         % (B * V * D * s / 1e6, B * V * Cr * s / 1e6, NSETS)},
        "roofline": ({"bound": "hbm", "kernel": "fused_fwd_kernel (score+bin+pool+fuse, one launch per step)",
                      "achieved": ach_fused, "peak": peak, "unit": "GB/s", "frac": ach_fused / peak, "traffic": traffic,
                      "peak_source": peak_src, "algorithmic_bytes_per_launch": ab["fwd"], "us_per_launch": t_fused * 1e3}
                     if fwd_launches == 1 else
                     {"bound": "hbm", "kernel": "pool_fuse_fwd_ring_kernel", "achieved": ach_pool, "peak": peak,
                      "unit": "GB/s", "frac": ach_pool / peak, "traffic": traffic, "peak_source": peak_src,
                      "algorithmic_bytes_per_launch": ab["pool_fwd"], "us_per_launch": t_pool * 1e3,
                      "timing": "CUDA event pair around each launch (includes ~4 us of event/launch gap); "
                                "back_to_back = K launches of this kernel in one event bracket",
                      "back_to_back": {"us_per_launch": ms_pool_b2b / K * 1e3,
                                       "achieved": ab["pool_fwd"] / (ms_pool_b2b / K * 1e-3) / 1e9,
                                       "frac": ab["pool_fwd"] / (ms_pool_b2b / K * 1e-3) / 1e9 / peak}}),
        "_roofline_rest": {
                     "other_kernels": {
                         "pool_fuse_fwd_ring_kernel": {"achieved": ach_pool, "frac": ach_pool / peak,
                                                       "us_per_launch": t_pool * 1e3, "algorithmic_bytes_per_launch": ab["pool_fwd"]},
                         "view_score_kernel": {"achieved": ach_score, "frac": ach_score / peak,
                                               "us_per_launch": t_score * 1e3, "algorithmic_bytes_per_launch": ab["score"]},
                         "pool_fuse_bwd_kernel": {"achieved": ach_bwd, "frac": ach_bwd / peak,
                                                  "us_per_launch": t_bwd * 1e3, "algorithmic_bytes_per_launch": ab["bwd"]},
                         "pool_fuse_fwd_kernel+mask": {"us_per_launch": t_pool_m * 1e3}},
                     "step": {"fwd_GBps": ach_fwd_step, "fwd_frac": ach_fwd_step / peak,
                              "fwd_bwd_GBps": ach_train_step, "fwd_bwd_frac": ach_train_step / peak}},
        "fwd_bwd": {"workload": "training step (BASELINE.json configs[2]): score+bin, pool+fuse with tie mask, "
                                "backward dS->dF, %s" % ("NCCL all-reduce of the %d-float FC-score gradient bucket "
                                                         "overlapped with the backward" % grad_bucket.numel()
                                                         if world > 1 else "no collective at N=1"),
                    "value": world * B * K / (ms_train * 1e-3), "unit": UNIT, "ms_per_step": ms_train / K,
                    "algorithmic_GBps_per_gpu": ach_train_step, "frac_of_peak": ach_train_step / peak,
                    "allreduce_alone_us": us_allreduce},
        "literal_batch_mode": {"workload": "same batch, reference-literal score_reduce='batch' (one scheme per batch, "
                                           "nets/model.py:146): 4 launches", "value": world * B * K / (ms_literal * 1e-3),
                               "unit": UNIT, "ms_per_step": ms_literal / K},
        "e2e": {"value": world * B * Ke / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": Ke, "ms_per_step": 1e3 * e2e_s / Ke,
                "api": "gvcnn_grouping_fusion_host (C ABI, pinned host buffers, chunk=%d shapes, 3-deep pipeline)" % chunk},
        "gpu_launches": fwd_launches * K,
        "clocks": clocks,
    }

    line["roofline"].update(line.pop("_roofline_rest"))
    if world == 1 and not args.no_cpu_baseline:
        Bc = B                                                       # the full configs[1] batch
        times, threads = cpu_reference_time(Bc, 48, 2)               # ~10-15 s of CPU work
        best = min(times)
        line["cpu_baseline"] = {"value": Bc / best, "unit": UNIT, "cores": threads, "kind": "port", "cpu_model": cpu_model(),
                                "mean_value": Bc * len(times) / sum(times),
                                "sample": "best of %d forward passes over the full %d-shape batch (%.0f ms best, %.1f s of "
                                          "CPU work): the reference's op graph (nets/model.py:16-102) restated in "
                                          "torch-CPU, all host threads" % (len(times), Bc, best * 1e3, sum(times))}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--e2e-chunk", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 4096 shapes per GPU (default, the contract's line); strong: 4096 shapes in total")
    ap.add_argument("--no-graph", action="store_true", help="time plain stream launches instead of a CUDA graph")
    args = ap.parse_args()
    if args.steps < 1:
        raise SystemExit("--steps must be >= 1")
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch ourselves under torchrun, one rank per GPU
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_cuda_arm(args)


if __name__ == "__main__":
    sys.exit(main())
