#!/usr/bin/env python
"""torchrun check of train_head.py on N GPUs: an epoch whose shards give the ranks DIFFERENT numbers of batches
(train_size 65, batch 4, 2 ranks: 9 / 8) must not hang, every rank must end with identical parameters (same
all-reduced gradients, same global-batch bins), and the loss must go down.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/train_head_multi_check.py
"""
import os
import sys
import tempfile

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("train_head", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                         "gvcnn-tf_b200", "train_head.py"))
th = importlib.util.module_from_spec(spec)
spec.loader.exec_module(th)

rank = int(os.environ["RANK"])
logdir = os.path.join(tempfile.gettempdir(), "gvcnn_train_multi_%s" % os.environ.get("MASTER_PORT", "0"))
orig_destroy = dist.destroy_process_group
dist.destroy_process_group = lambda *a, **k: None                  # keep the group for the comparison below
hist = th.main(["--how_many_training_epochs", "3", "--train_size", "65", "--val_size", "16", "--batch_size", "4",
                "--num_views", "12", "--raw_channels", "1024", "--final_channels", "1024", "--feature_hw", "2",
                "--base_learning_rate", "0.01", "--train_logdir", logdir, "--score_reduce", sys.argv[1] if len(sys.argv) > 1 else "batch"])
dist.barrier()
# parameters after training must be identical on every rank (same all-reduced gradients, same global-batch bins)
flat = torch.cat([p.detach().reshape(-1).float() for p in th.main.last_head.parameters()])
gathered = [torch.zeros_like(flat) for _ in range(dist.get_world_size())]
dist.all_gather(gathered, flat)
assert all(torch.equal(g, gathered[0]) for g in gathered), "parameters diverged across ranks"
losses = [h[1] for h in hist]
if rank == 0:
    print("epochs", [(e, round(l, 4), round(a, 3)) for e, l, a in hist])
    assert losses[-1] < losses[0], losses
    print("train_head multi-GPU check ok: world", dist.get_world_size())
orig_destroy()
