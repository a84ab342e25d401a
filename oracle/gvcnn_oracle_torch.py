"""Graph-literal torch-CPU restatement of nets/model.py:44-102.  TEST INFRASTRUCTURE ONLY.

Two uses, both as the checker / the timed CPU baseline, never as product code:
  * gradients: torch autograd over the same op graph TF differentiates
    (``torch.amax`` splits the gradient equally among ties exactly like TF's
    ``_MinOrMaxGrad``; ``torch.max(dim)`` would not - never use it here);
  * ``bench.py``'s ``cpu_baseline`` / ``--impl reference``: the reference's op
    SEQUENCE (stack all V views -> per group: where -> gather or ones dummy ->
    reduce -> multiply by w_g -> add_n -> divide) run multi-threaded on the
    host, one materialised tensor per op like a TF1 graph executes it.

Parity unpinned against live TensorFlow (not installable here); pinned by the
KATs in tests/golden/ and by agreement with oracle/gvcnn_oracle.py.
"""
from __future__ import annotations

import torch


def view_pooling(final_view_descriptors, group_scheme, pool="max", empty_fill=1.0):
    """nets/model.py:44-74 op for op.  ``final_view_descriptors`` is a list of V
    tensors; ``tf.ones_like(list)`` and ``tf.gather(list, ind)`` both pack the
    list into a [V, ...] tensor first, which is why the stack is repeated per
    group here (TF1 does not CSE the implicit pack ops created by separate
    convert_to_tensor calls)."""
    group_descriptors = {}
    dummy = torch.full_like(torch.stack(list(final_view_descriptors)), empty_fill)
    scheme_list = torch.unbind(torch.as_tensor(group_scheme))
    indices = [torch.nonzero(elem, as_tuple=False).squeeze(1) for elem in scheme_list]
    for i, ind in enumerate(indices):
        if ind.numel() > 0:
            pooled_view = torch.index_select(torch.stack(list(final_view_descriptors)), 0, ind)
        else:
            pooled_view = dummy
        if pool == "max":
            group_descriptors[i] = torch.amax(pooled_view, dim=0)
        else:
            group_descriptors[i] = torch.sum(pooled_view, dim=0) / pooled_view.shape[0]
    return group_descriptors


def group_fusion(group_descriptors, group_weight):
    """nets/model.py:77-102 op for op."""
    group_weight_list = torch.unbind(torch.as_tensor(group_weight, dtype=torch.float32))
    numerator = []
    for key, value in group_descriptors.items():
        numerator.append(group_weight_list[key] * value)
    denominator = torch.stack(group_weight_list).sum()
    acc = numerator[0]
    for t in numerator[1:]:
        acc = acc + t
    return acc / denominator


def view_scores(R, W, b, score_reduce="shape"):
    """nets/model.py:144-147: per-view Dense(1) on the post-GAP raw descriptor,
    optional batch mean, sigmoid(log|x|).  R [B,V,C], W [V,C], b [V]."""
    V = R.shape[1]
    xs = []
    for v in range(V):
        raw = R[:, v, :] @ W[v][:, None] + b[v]          # Dense(1) -> [B, 1]
        if score_reduce == "batch":
            raw = raw.mean()
        else:
            raw = raw[:, 0]
        xs.append(raw)
    x = torch.stack(xs, dim=-1)
    return x, torch.sigmoid(torch.log(torch.abs(x)))


def scheme_from_bins(brow, num_group):
    V = len(brow)
    scheme = torch.zeros((num_group, V), dtype=torch.int64)
    scheme[torch.as_tensor(brow, dtype=torch.int64), torch.arange(V)] = 1
    return scheme


def pool_fuse(F, bins, num_group, pool="max", empty_fill=1.0):
    """Per-shape-bin-map forward, differentiable.  F [B,V,D]; bins [B,V] or [V].
    Shapes sharing a bin map are run through one literal graph."""
    B, V, D = F.shape
    bins = torch.as_tensor(bins)
    if bins.dim() == 1:
        bins = bins[None, :].expand(B, V)
    uniq, inv = torch.unique(bins, dim=0, return_inverse=True)
    out = [None] * len(uniq)
    S = torch.zeros((B, D), dtype=F.dtype)
    pieces, index = [], []
    for u in range(len(uniq)):
        sel = torch.nonzero(inv == u).squeeze(1)
        scheme = scheme_from_bins(uniq[u], num_group)
        w = 1.0 + scheme.sum(dim=1).to(torch.float32)
        views = [F[sel, v, :] for v in range(V)]
        desc = view_pooling(views, scheme, pool=pool, empty_fill=empty_fill)
        pieces.append(group_fusion(desc, w))
        index.append(sel)
    S = torch.zeros((B, D), dtype=F.dtype).index_copy(0, torch.cat(index), torch.cat(pieces))
    return S


def reference_step_cpu(F_views, R, W, b, num_group, pool="max", empty_fill=1.0):
    """One reference-shaped forward of the path on the host for timing, in the
    reference's real call pattern (train.py:270-288): scores -> host binning
    (per-batch scheme, model.py:16-41) -> literal pooling + fusion graph on the
    V view tensors.  Returns S [N, D]."""
    x, s = view_scores(R, W, b, score_reduce="batch")
    G = num_group
    V = len(F_views)
    scheme = torch.zeros((G, V), dtype=torch.int64)
    for idx, score in enumerate(s.reshape(-1).tolist()):
        # nets/model.py:23 - an out-of-range bin raises IndexError (score == 1.0), a NaN score ValueError; no clamp
        scheme[int(torch.tensor(score, dtype=torch.float32) * G), idx] = 1
    w = torch.zeros(G, dtype=torch.float32)
    for i in range(G):
        w[i] = 1 + int(scheme[i].sum())
    desc = view_pooling(F_views, scheme, pool=pool, empty_fill=empty_fill)
    return group_fusion(desc, w)


def paper_mode(R, W, b, F, num_group, pool="max"):
    """Differentiable paper-mode path (no reference counterpart), for autograd checks:
    x = R.W + b; s = |x| / (1 + |x|); bins = trunc(float32(s) * G) (not differentiated);
    w_g = mean score of the group's views (0 if empty); S = sum_g w_g P_g / sum_g w_g.
    R [B,V,C], F [B,V,D]; any float dtype (use float64).  Returns (S, s, bins, w)."""
    B, V, _ = R.shape
    x = torch.einsum("bvc,vc->bv", R, W) + b[None, :]
    s = torch.abs(x) / (1 + torch.abs(x))
    bins = torch.trunc(s.detach().to(torch.float32) * num_group).to(torch.int64)
    onehot = torch.nn.functional.one_hot(bins, num_group).to(R.dtype)          # [B, V, G]
    cnt = onehot.sum(dim=1)                                                    # [B, G]
    w = torch.where(cnt > 0, (onehot * s[:, :, None]).sum(dim=1) / cnt.clamp(min=1), torch.zeros_like(cnt))
    big = torch.finfo(F.dtype).max
    Ps = []
    for g in range(num_group):
        member = onehot[:, :, g] > 0                                           # [B, V]
        if pool == "max":
            Pg = torch.amax(torch.where(member[:, :, None], F, torch.full_like(F, -big)), dim=1)
        else:
            Pg = (F * member[:, :, None]).sum(dim=1) / cnt[:, g].clamp(min=1)[:, None]
        Ps.append(torch.where((cnt[:, g] > 0)[:, None], Pg, torch.zeros_like(Pg)))
    P = torch.stack(Ps, dim=1)                                                 # [B, G, D]
    S = (w[:, :, None] * P).sum(dim=1) / w.sum(dim=1, keepdim=True)
    return S, s, bins, w
