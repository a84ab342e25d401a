"""CPU oracle for the GVCNN view-grouping + fusion path.  TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``gvcnn-tf_b200/`` imports it.

It restates, op for op and in the reference's own op ORDER, what
``/root/reference/nets/model.py`` computes between the per-view backbone
descriptors and the classifier.  Every function cites the reference lines it
follows.  All float32 work is done with NumPy float32 array ops, which are IEEE
one-rounding-per-op (no FMA contraction), so the result of every expression
below is a well defined bit pattern that the CUDA kernels are required to
reproduce exactly (they use __fmul_rn/__fadd_rn/__fdiv_rn in the same order).

Parity status: PARITY UNPINNED against a live TensorFlow run — TensorFlow 1.x
is not installable in this image (no wheel, no network, Python 3.12) and the
reference records no expected outputs.  What pins this oracle instead:
  * KAT-1 / KAT-2 / identity KAT hand-derived from ``unit_test.py:18-19`` and
    ``nets/model.py`` (tests/golden/kat.json),
  * the reference's OWN ``group_scheme`` / ``group_weight`` NumPy code and its
    OWN ``view_pooling`` / ``group_fusion`` graph-construction code, executed
    here in the container over a NumPy stand-in for the 13 TF ops they call
    (tests/golden/make_golden.py -> tests/golden/ref_graph_*.npz),
  * a second, independent C restatement (oracle/gvcnn_oracle.c) that must
    agree bit for bit.

Mode flags:
  pool        'max'  -> nets/model.py:72 (shipped)   | 'mean' -> unit_test.py:31
  empty_fill  1.0    -> nets/model.py:63 (ones dummy) | 0.0   -> unit_test.py:22
  score_reduce 'batch' -> nets/model.py:146 (reduce_mean over the batch, one
               scheme shared by the batch) | 'shape' -> the same code run at
               batch size 1 per shape (bin map [B, V]).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "group_scheme", "group_weight", "view_pooling", "group_fusion",
    "view_scores", "score_from_x", "bins_from_scores", "edge_ulps_distance",
    "pool_fuse_fwd", "pool_fuse_bwd", "grouping_fusion_fwd", "round_bf16",
    "sorted_view_order", "tie_mask_planes", "gap_mean_kernel_order", "add_n", "order_edge",
]


# --------------------------------------------------------------------------
# host part: binning and weights                       nets/model.py:16-41
# --------------------------------------------------------------------------
def group_scheme(view_discrimination_score, num_group, num_views, multiplier=None):
    """One-hot [num_group, num_views] scheme.  Follows nets/model.py:16-25.

    ``view_discrimination_score`` is what ``sess.partial_run(h, [view_scores])``
    returns (train.py:270-276): a list holding ONE list of V float32 scalars,
    hence the ``[0]`` below (model.py:22).  The reference multiplies by a
    hard-coded 10 (model.py:23) and is only ever run with num_group == 10
    (train.py:97, eval.py:23); ``multiplier=None`` generalises that to
    ``num_group`` (identical at 10), ``multiplier=10`` is the literal code.

    float32(score) * float32(multiplier) is rounded to float32 BEFORE int()
    truncates (NumPy scalar arithmetic of an np.float32 with a Python int).
    Errors mirror the reference: IndexError when the bin is >= num_group
    (score == 1.0, i.e. |x| >= 2**24), ValueError for a NaN score.
    """
    if multiplier is None:
        multiplier = num_group
    schemes = np.full((num_group, num_views), 0, dtype=np.int64)
    for idx, score in enumerate(view_discrimination_score[0]):
        t = np.float32(score) * np.float32(multiplier)
        if np.isnan(t):
            raise ValueError("cannot convert float NaN to integer")
        b = int(t)
        if b >= num_group or b < -num_group:
            raise IndexError(
                "index %d is out of bounds for axis 0 with size %d" % (b, num_group))
        schemes[b, idx] = 1
    return schemes


def group_weight(g_schemes):
    """w[g] = 1 + number of views in group g.  Follows nets/model.py:28-41
    (``sum = 1`` at :34, float32 result at :32)."""
    g_schemes = np.asarray(g_schemes)
    num_group, num_views = g_schemes.shape
    weights = np.zeros(shape=(num_group,), dtype=np.float32)
    for i in range(num_group):
        s = 1
        for j in range(num_views):
            if g_schemes[i][j] == 1:
                s += g_schemes[i][j]
        weights[i] = s
    return weights


# --------------------------------------------------------------------------
# graph part, literal op sequence                      nets/model.py:44-102
# --------------------------------------------------------------------------
def _reduce_mean_axis0(x):
    """tf.reduce_mean(x, axis=0): sequential sum over the leading axis, then
    one division by the count.  Integer dtypes divide truncating toward zero
    (C++ integer division inside the Eigen MeanReducer) - that is what makes
    KAT-1's g3 = [1, 10, 85, 10]."""
    acc = x[0].copy()
    for j in range(1, x.shape[0]):
        acc = acc + x[j]
    n = x.shape[0]
    if np.issubdtype(x.dtype, np.integer):
        return (np.trunc(acc.astype(np.float64) / n)).astype(x.dtype)
    return (acc / x.dtype.type(n)).astype(x.dtype)


def view_pooling(final_view_descriptors, group_scheme, pool="max", empty_fill=1.0):
    """dict{g: pooled descriptor}.  Follows nets/model.py:44-74.

    final_view_descriptors: list of V arrays of identical shape (the implicit
    tf.stack at model.py:63/69).  For every group: tf.where -> indices; a
    non-empty group gathers its views, an empty one takes the dummy
    (model.py:68-70, ones at :63; zeros in unit_test.py:22); then reduce over
    axis 0 (max at model.py:72, mean at unit_test.py:31).
    """
    stacked = np.stack([np.asarray(f) for f in final_view_descriptors])
    dummy = np.full_like(stacked, empty_fill)
    group_descriptors = {}
    for i, elem in enumerate(np.asarray(group_scheme)):
        ind = np.where(elem)[0]
        pooled_view = stacked[ind] if ind.size > 0 else dummy
        if pool == "max":
            group_descriptors[i] = pooled_view.max(axis=0)
        elif pool == "mean":
            group_descriptors[i] = _reduce_mean_axis0(pooled_view)
        else:
            raise ValueError(pool)
    return group_descriptors


def add_n(terms, association="left"):
    """tf.add_n(terms) (nets/model.py:100), one float32 rounding per addition, in one of two association orders:

    'left'  ((t0 + t1) + t2) + ...   - the op-order MODEL the CUDA kernels implement and are held bit-exact to;
    'tf8'   TensorFlow's CPU AddN kernel as remembered from tensorflow/core/kernels/aggregate_ops.cc (NOT verifiable
            here - no TensorFlow in this image): the first r = N % 8 inputs (r = 8 when N % 8 == 0, r = 9 when
            N % 8 == 1) are added left to right into the output, then every further block of 8 inputs is summed
            left to right and added to the output as ONE term:  out = out + (((t_r + t_r+1) + ...) + t_r+7).
            N = 10 (the reference's only num_group): (t0 + t1) + (t2 + ... + t9).  Identical to 'left' for N <= 9.
    The two differ in the last bits only; tests hold the CUDA result within the north star's 1e-5 of BOTH."""
    terms = list(terms)
    n = len(terms)
    if association == "left" or n <= 9:
        acc = terms[0]
        for t in terms[1:]:
            acc = acc + t
        return acc
    if association != "tf8":
        raise ValueError(association)
    r = n % 8
    if r == 0:
        r = 8
    elif r == 1:
        r = 9
    acc = terms[0]
    for t in terms[1:r]:
        acc = acc + t
    while r < n:
        blk = terms[r]
        for t in terms[r + 1:r + 8]:
            blk = blk + t
        acc = acc + blk
        r += 8
    return acc


def group_fusion(group_descriptors, group_weight, association="left"):
    """S = add_n_g(w_g * P_g) / sum_g w_g.  Follows nets/model.py:77-102:
    multiply per group (:97), reduce_sum of the weights (:99), add_n in dict
    order 0..G-1 (``association``: see add_n), one true division (:100)."""
    w = np.asarray(group_weight, dtype=np.float32)
    numerator = [w[key] * value for key, value in group_descriptors.items()]
    denominator = np.float32(0)
    for x in w:
        denominator = np.float32(denominator + x)
    return add_n(numerator, association) / denominator


# --------------------------------------------------------------------------
# score                                                nets/model.py:143-148
# --------------------------------------------------------------------------
def view_scores(R, W, b, score_reduce="shape", dtype=np.float64):
    """Raw FC output x and score s per view.  Follows nets/model.py:144-147.

    R [B, V, C] is the post-GAP raw descriptor (model.py:144), W [V, C] / b [V]
    are the V separate Dense(1) layers (model.py:145 is inside the view loop),
    x = R_v . W_v + b_v; 'batch' then takes reduce_mean over the batch
    (model.py:146) giving one scalar per view, 'shape' keeps [B, V].
    ``dtype`` float64 gives the reference value of the mathematics (TF's own
    float32 summation order inside Eigen is not reproducible without TF);
    float32 gives a plain left-to-right float32 evaluation.
    Returns (x, s) with s = sigmoid(log(|x|)) computed in ``dtype``.
    """
    R = np.asarray(R, dtype=dtype)
    W = np.asarray(W, dtype=dtype)
    b = np.asarray(b, dtype=dtype)
    if dtype == np.float64:
        x = np.einsum("bvc,vc->bv", R, W) + b[None, :]
    else:
        x = np.zeros(R.shape[:2], dtype=dtype)
        for c in range(R.shape[2]):
            x = x + R[:, :, c] * W[None, :, c]
        x = x + b[None, :]
    if score_reduce == "batch":
        acc = x[0].copy()
        for j in range(1, x.shape[0]):
            acc = acc + x[j]
        x = (acc / dtype(x.shape[0]))[None, :]
    elif score_reduce != "shape":
        raise ValueError(score_reduce)
    return x, score_from_x(x)


def gap_mean_kernel_order(maps, nslices=8):
    """GlobalAveragePooling2D of channel-last maps (nets/model.py:144: [N, h, w, C] -> [N, C]) = tf.reduce_mean over
    the positions: a float32 sum, then ONE division by the count.  TF's own summation order (Eigen's tree) is not
    reproducible without TF, so the order restated here is the CUDA kernel's (csrc/gap_score.cu): the HW positions are
    cut into `nslices` contiguous slices (the first HW % nslices have one more), each summed sequentially from 0.0, the
    slice sums added in ascending order, then divided by HW.  maps [..., HW, C] float32 -> [..., C] float32."""
    m = np.asarray(maps, dtype=np.float32)
    HW = m.shape[-2]
    q, r = divmod(HW, nslices)
    parts = []
    p0 = 0
    for s_ in range(nslices):
        n = q + (1 if s_ < r else 0)
        acc = np.zeros(m.shape[:-2] + m.shape[-1:], dtype=np.float32)
        for p in range(p0, p0 + n):
            acc = acc + m[..., p, :]
        parts.append(acc)
        p0 += n
    tot = parts[0]
    for a in parts[1:]:
        tot = tot + a
    return (tot / np.float32(HW)).astype(np.float32)


def score_from_x(x):
    """s = tf.nn.sigmoid(tf.math.log(tf.abs(x))) in x's dtype (model.py:147).
    log(0) = -inf -> s = 0;  |x| = inf -> s = 1;  NaN stays NaN."""
    x = np.asarray(x)
    with np.errstate(divide="ignore", over="ignore", invalid="ignore"):
        y = np.log(np.abs(x))
        return (1 / (1 + np.exp(-y))).astype(x.dtype)


def score_from_x_rational(x):
    """|x| / (1 + |x|): the same function of x, one division; what the CUDA
    kernel evaluates (float32, __fdiv_rn).  inf -> 1."""
    x = np.asarray(x)
    ax = np.abs(x)
    with np.errstate(invalid="ignore"):
        s = ax / (1 + ax)
    return np.where(np.isinf(ax), x.dtype.type(1), s).astype(x.dtype)


def bins_from_scores(s, num_group, multiplier=None):
    """Vectorised model.py:23: int(float32(s) * float32(multiplier)).
    Returns int32 bins with no range check (bin == num_group marks the
    reference's IndexError case; NaN -> INT32_MIN marks its ValueError)."""
    if multiplier is None:
        multiplier = num_group
    t = np.asarray(s, dtype=np.float32) * np.float32(multiplier)
    out = np.full(t.shape, np.iinfo(np.int32).min, dtype=np.int32)
    ok = ~np.isnan(t)
    out[ok] = np.trunc(t[ok]).astype(np.int32)
    return out


def edge_ulps_distance(s, num_group, k=1):
    """True where moving the float32 score by up to k ulps changes its bin -
    north_star's "score within 1 ulp of a bin edge" (reported separately)."""
    s = np.asarray(s, dtype=np.float32)
    lo, hi = s.copy(), s.copy()
    for _ in range(k):
        lo = np.nextafter(lo, np.float32(-np.inf))
        hi = np.nextafter(hi, np.float32(np.inf))
    b = bins_from_scores(s, num_group)
    return (bins_from_scores(lo, num_group) != b) | (bins_from_scores(hi, num_group) != b)


def order_edge(xm, A, n_terms, num_group, edge_ulps=1, multiplier=None):
    """The a-priori order-sensitivity report (GVCNN_FLAG_ORDER_EDGE; csrc/common.cuh order_edge_flag), float32 op
    for op: two rounded evaluations of x = sum of n_terms products lie within dx = 2 gamma_n A of each other
    (A = sum |r_c w_c| + |b|, gamma_n = n u / (1 - n u), u = 2^-24); True where [s(|x| - dx), s(|x| + dx)], widened
    by edge_ulps ulps, straddles a bin edge - i.e. where another float32 evaluation order (TensorFlow's) may
    legitimately produce a different group index."""
    f = np.float32
    xm = np.asarray(xm, dtype=np.float32)
    A = np.asarray(A, dtype=np.float32)
    nu = f(n_terms) * f(5.9604645e-8)
    dx = (f(2.0) * (nu / (f(1.0) - nu))) * A
    ax = np.abs(xm)
    with np.errstate(invalid="ignore", over="ignore"):
        ok = (ax < f(3.0e38)) & (dx < f(3.0e38))
        lo = np.maximum(ax - dx, f(0.0))
        hi = ax + dx
        s_lo = (lo / (f(1.0) + lo)).astype(np.float32)
        s_hi = (hi / (f(1.0) + hi)).astype(np.float32)
    bl = s_lo.view(np.uint32).astype(np.int64)
    bh = s_hi.view(np.uint32).astype(np.int64)
    s_lo = np.where(bl > edge_ulps, bl - edge_ulps, 0).astype(np.uint32).view(np.float32)
    s_hi = (bh + edge_ulps).astype(np.uint32).view(np.float32)
    fg = f(multiplier if multiplier else num_group)
    with np.errstate(invalid="ignore"):
        out = np.trunc(s_lo * fg) != np.trunc(s_hi * fg)
    return out & ok


# --------------------------------------------------------------------------
# batched restatement used at scale (same arithmetic, same order)
# --------------------------------------------------------------------------
def round_bf16(x):
    """float32 -> nearest-even bfloat16, returned as float32."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    out = r.astype(np.uint32).view(np.float32)
    return np.where(np.isnan(x), x, out)


def pool_fuse_fwd(F, bins, num_group, pool="max", empty_fill=1.0, association="left"):
    """Pooling + fusion for a whole batch with per-shape bin maps.

    F [B, V, D] float32, bins [B, V] (or [V], shared by the batch).  Runs the
    literal per-shape graph of view_pooling + group_fusion (batch size 1 per
    shape shares no arithmetic across shapes, so this is exactly the reference
    at batch 1) but vectorised over shapes that share a bin map.
    Returns S [B, D] float32.
    """
    F = np.asarray(F, dtype=np.float32)
    B, V, D = F.shape
    bins = np.asarray(bins)
    if bins.ndim == 1:
        bins = np.broadcast_to(bins[None, :], (B, V))
    S = np.empty((B, D), dtype=np.float32)
    uniq, inv = np.unique(bins, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    for u, brow in enumerate(uniq):
        sel = np.where(inv == u)[0]
        scheme = np.zeros((num_group, V), dtype=np.int64)
        scheme[brow, np.arange(V)] = 1
        w = group_weight(scheme)
        views = [F[sel, v, :] for v in range(V)]
        desc = view_pooling(views, scheme, pool=pool, empty_fill=empty_fill)
        S[sel] = group_fusion(desc, w, association)
    return S


def sorted_view_order(brow):
    """Views ordered by (bin, view index) - the deterministic order both CUDA
    kernels derive; tie-mask bit k refers to the k-th view in this order."""
    brow = np.asarray(brow)
    return np.lexsort((np.arange(brow.size), brow))


def pool_fuse_bwd(dS, F, bins, num_group, pool="max", weights=None):
    """dF [B, V, D] from dS [B, D].  Follows TF autodiff of model.py:62-100
    (SURVEY.md section 3.4), in TF's op order:
      div grad      g0 = dS / sum_w                     (_RealDivGrad)
      add_n grad    passes g0 to every term
      multiply grad g1 = g0 * w_g
      reduce_max    dF_v = (indicator / num_selected) * g1   (_MinOrMaxGrad:
                    ties share equally)
      reduce_mean   dF_v = g1 / n_g                      (_MeanGrad)
      cond/gather   only the taken branch; empty groups' dummy gets no grad;
                    every view is in exactly one group so the scatter-add of
                    the gather grads is a plain assignment.
    No gradient reaches scores / W / b / weights (train.py:127-128 feeds them
    through placeholders; utils/train_utils.py:203-206 skips None grads).
    """
    dS = np.asarray(dS, dtype=np.float32)
    F = np.asarray(F, dtype=np.float32)
    B, V, D = F.shape
    bins = np.asarray(bins)
    if bins.ndim == 1:
        bins = np.broadcast_to(bins[None, :], (B, V))
    dF = np.zeros_like(F)
    uniq, inv = np.unique(bins, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    sumw = np.float32(num_group + V)
    if weights is not None:                       # caller-supplied weights [G] (group_fusion's second argument)
        weights = np.asarray(weights, dtype=np.float32)
        sumw = np.float32(0)
        for x in weights:
            sumw = np.float32(sumw + x)
    for u, brow in enumerate(uniq):
        sel = np.where(inv == u)[0]
        g0 = dS[sel] / sumw
        for g in range(num_group):
            ind = np.where(brow == g)[0]
            if ind.size == 0:
                continue
            g1 = g0 * (np.float32(1 + ind.size) if weights is None else weights[g])
            if pool == "max":
                x = F[sel][:, ind, :]                       # [n, k, D]
                y = x.max(axis=1, keepdims=True)
                indicators = (x == y).astype(np.float32)
                num_selected = indicators.sum(axis=1, keepdims=True, dtype=np.float32)
                dF[np.ix_(sel, ind)] = (indicators / num_selected) * g1[:, None, :]
            else:
                dF[np.ix_(sel, ind)] = (g1 / np.float32(ind.size))[:, None, :]
    return dF


def tie_mask_planes(F, bins, num_group):
    """Max-mode routing aid the forward kernel saves for the backward:
    uint8 planes [ceil(V/8), B, D]; bit (k % 8) of plane k // 8 is set iff the
    k-th view in sorted_view_order attains its group's maximum."""
    F = np.asarray(F, dtype=np.float32)
    B, V, D = F.shape
    bins = np.asarray(bins)
    if bins.ndim == 1:
        bins = np.broadcast_to(bins[None, :], (B, V))
    P = (V + 7) // 8
    out = np.zeros((P, B, D), dtype=np.uint8)
    for bidx in range(B):
        order = sorted_view_order(bins[bidx])
        for k, v in enumerate(order):
            grp = np.where(bins[bidx] == bins[bidx, v])[0]
            m = F[bidx, grp, :].max(axis=0)
            out[k // 8, bidx] |= ((F[bidx, v] == m).astype(np.uint8) << (k % 8))
    return out


def grouping_fusion_fwd(R, W, b, F, num_group, pool="max", empty_fill=1.0,
                        score_reduce="shape", score_dtype=np.float64):
    """Whole forward path: scores -> bins -> pooled/fused descriptor.
    Returns dict(x, scores, bins, S).  Raises like group_scheme does."""
    x, s = view_scores(R, W, b, score_reduce=score_reduce, dtype=score_dtype)
    s32 = s.astype(np.float32)
    bins = bins_from_scores(s32, num_group)
    if np.any(bins == np.iinfo(np.int32).min):
        raise ValueError("cannot convert float NaN to integer")
    if np.any(bins >= num_group):
        raise IndexError("index %d is out of bounds for axis 0 with size %d"
                         % (int(bins.max()), num_group))
    bmap = bins[0] if score_reduce == "batch" else bins
    S = pool_fuse_fwd(F, bmap, num_group, pool=pool, empty_fill=empty_fill)
    return {"x": x, "scores": s32, "bins": bins, "S": S}
