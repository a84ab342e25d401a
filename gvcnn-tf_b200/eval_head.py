#!/usr/bin/env python
"""eval.py-shaped driver for the grouping + fusion head.

Mirrors the reference's ``eval.py`` for the part of the graph this repo owns:

    reference (eval.py:110-215)                        here
    ------------------------------------------------  ------------------------------------------------
    Saver.restore(latest checkpoint in a directory,    torch.load of the newest ``<ckpt_name>-NNNN`` in
    or the given file)                  (:120-126)     --checkpoint_path, or of the given file
    per batch: partial_run #1 -> host group_scheme /   one device pass per batch: score+bin -> pool+fuse
    group_weight -> partial_run #2      (:177-202)     -> GAP -> Dense(num_classes), eval mode, no host hop
    accuracy = mean over batches of the batch          the same figure (and the per-shape accuracy beside
    accuracy; summed confusion matrix   (:204-215)     it: they differ when the last batch is short)

The per-view features come from a feature shard (``records.FeatureShard`` prefix in --dataset_path, the
pre-extracted-feature format of SURVEY 8f n4) or, if no shard is there, from the seeded synthetic source
of train_head.py (its validation split).  The head geometry is read from the checkpoint.
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvcnn_tf_b200 import model, records  # noqa: E402
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("_gvcnn_train_head", os.path.join(os.path.dirname(os.path.abspath(__file__)), "train_head.py"))
train_head = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(train_head)


def build_flags():
    """Flag names and defaults of eval.py:30-43 (num_group: eval.py:63 reads FLAGS.num_group, NUM_GROUP = 10)."""
    p = argparse.ArgumentParser()
    p.add_argument("--dataset_path", default="/home/ace19/dl_data/modelnet/test.record")
    p.add_argument("--checkpoint_path", default=os.path.join(os.getcwd(), "models"))
    p.add_argument("--batch_size", type=int, default=4)
    p.add_argument("--num_views", type=int, default=6)
    p.add_argument("--height", type=int, default=299)          # accepted for drop-in command lines; the backbone
    p.add_argument("--width", type=int, default=299)           # that consumes them is out of scope
    p.add_argument("--labels", default="airplane,bed,bookshelf,toilet,vase")
    p.add_argument("--num_group", type=int, default=10)
    p.add_argument("--ckpt_name", default="gvcnn.ckpt")
    return p


def latest_checkpoint(path, name):
    """tf.train.latest_checkpoint for the files train_head.py writes (eval.py:121-125)."""
    if os.path.isdir(path):
        cks = sorted(f for f in os.listdir(path) if f.startswith(name))
        if not cks:
            raise FileNotFoundError("no %s-* checkpoint in %s" % (name, path))
        return os.path.join(path, cks[-1])
    return path


def main(argv=None):
    flags = build_flags().parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("eval_head.py needs a CUDA device: the grouping/fusion path has no CPU fallback")
    device = torch.device("cuda", 0)
    labels = flags.labels.split(",")
    num_classes = len(labels)
    ck = torch.load(latest_checkpoint(flags.checkpoint_path, flags.ckpt_name), map_location=device)
    tf = argparse.Namespace(**ck["flags"])                      # the geometry the head was trained with
    if (tf.num_views, tf.num_group, len(tf.labels.split(","))) != (flags.num_views, flags.num_group, num_classes):
        raise ValueError("checkpoint was trained with num_views=%d num_group=%d and %d labels"
                         % (tf.num_views, tf.num_group, len(tf.labels.split(","))))
    head = model.GVCNNHead(tf.num_views, tf.raw_channels, tf.final_channels, num_classes, num_group=tf.num_group,
                           score_reduce=tf.score_reduce, weight_mode=tf.weight_mode).to(device)
    head.load_state_dict(ck["head"])
    head.eval()

    if os.path.exists(flags.dataset_path + ".raw.npy"):
        shard = records.FeatureShard(flags.dataset_path)
        n_total = len(shard)
        batches = ((r.to(device, non_blocking=True), f.to(device, non_blocking=True), y.to(device))
                   for r, f, y in shard.batches(flags.batch_size))
    else:
        src = train_head.SyntheticFeatures(tf.val_size, tf, num_classes, tf.seed + 2, device)
        n_total = tf.val_size
        batches = src.batches(flags.batch_size)

    count, total_acc, correct = 0, 0.0, 0
    cm = torch.zeros((num_classes, num_classes), dtype=torch.int64)
    with torch.no_grad():
        for raw, final, y in batches:
            _, _, logits = head(raw.float(), final.float())
            pred = logits.argmax(dim=1)
            hit = int((pred == y).sum())
            total_acc += hit / len(y)                           # eval.py:204: the batch's accuracy
            correct += hit
            count += 1
            cm += train_head.confusion_matrix(y.cpu(), pred.cpu(), num_classes)
    total_acc /= max(count, 1)
    print("Confusion Matrix:\n %s" % cm)
    print("Final test accuracy = %.3f%% (N=%d)" % (total_acc * 100, n_total))
    return {"accuracy": total_acc, "per_shape_accuracy": correct / max(n_total, 1), "confusion_matrix": cm, "n": n_total}


if __name__ == "__main__":
    main()
