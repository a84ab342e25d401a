"""The oracle against every fixture the reference offers for this path (CPU only).

tests/golden/kat.json          hand-derived known answers on unit_test.py:18-19
tests/golden/ref_graph_*.npz   the reference's own group_scheme / group_weight /
                               view_pooling / group_fusion code run in the build
                               container (tests/golden/make_golden.py)
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import gvcnn_oracle as O
from oracle import gvcnn_oracle_torch as OT


@pytest.fixture(scope="module")
def kat(golden_dir):
    with open(os.path.join(golden_dir, "kat.json")) as f:
        return json.load(f)


def test_kat1_mean_zeros_int32(kat):
    """unit_test.py:18-31: reduce_mean on int32 truncates; empty groups are zeros."""
    F = np.array(kat["F"], dtype=np.int32)
    desc = O.view_pooling([f for f in F], np.array(kat["scheme"]), pool="mean", empty_fill=0)
    for g, want in kat["kat1_mean_zeros_int32"].items():
        assert desc[int(g)].tolist() == want
    descf = O.view_pooling([f for f in F.astype(np.float32)], np.array(kat["scheme"]), pool="mean", empty_fill=0.0)
    np.testing.assert_allclose(descf[3], kat["kat1_mean_zeros_float_g3"], rtol=1e-6)


def test_kat2_model_literal(kat):
    """nets/model.py:28-41,62-100 on the same data: max pooling, ones for empty groups."""
    F = np.array(kat["F"], dtype=np.float32)
    scheme = np.array(kat["scheme"])
    w = O.group_weight(scheme)
    assert w.tolist() == kat["kat2_weights"] and w.dtype == np.float32
    desc = O.view_pooling([f for f in F], scheme)
    for g, want in kat["kat2_max_ones_groups"].items():
        assert desc[int(g)].tolist() == want
    S = O.group_fusion(desc, w)
    assert S.dtype == np.float32
    np.testing.assert_array_equal(S, np.array(kat["kat2_S"], dtype=np.float32))
    # the batched restatement and the C restatement give the same bits
    S2 = O.pool_fuse_fwd(F[None], np.array(kat["bins"]), 5)
    np.testing.assert_array_equal(S2[0], S)


def test_kat2_c_oracle(kat, c_oracle):
    F = np.array(kat["F"], dtype=np.float32)
    S = c_oracle.pool_fuse_fwd(F[None], np.array(kat["bins"], dtype=np.int32)[None], 5)
    np.testing.assert_array_equal(S[0], np.array(kat["kat2_S"], dtype=np.float32))


def test_identity_kat():
    """All views in one group g: P_g = basic's max over all views (nets/model.py:202),
    S = ((1+V) * P_g + (G-1) * 1) / (G+V)."""
    rng = np.random.default_rng(1)
    V, G, D = 6, 10, 32
    F = rng.standard_normal((3, V, D)).astype(np.float32)
    for g in (0, 4, 9):
        S = O.pool_fuse_fwd(F, np.full(V, g), G)
        want = (np.float32(1 + V) * F.max(axis=1) + np.float32(G - 1)) / np.float32(G + V)
        np.testing.assert_allclose(S, want, rtol=2e-6)


def test_reference_host_functions(golden_dir):
    """group_scheme / group_weight of the reference (run unmodified) == oracle."""
    z = np.load(os.path.join(golden_dir, "ref_graph_host.npz"))
    for i in range(int(z["n"])):
        sc = z["scores_%d" % i]
        scheme = O.group_scheme([list(sc)], 10, len(sc))
        np.testing.assert_array_equal(scheme, z["scheme_%d" % i])
        np.testing.assert_array_equal(O.group_weight(scheme), z["weight_%d" % i])
        # literal multiplier 10 == generalised multiplier at num_group 10
        np.testing.assert_array_equal(O.group_scheme([list(sc)], 10, len(sc), multiplier=10), scheme)
        bins = O.bins_from_scores(sc, 10)
        np.testing.assert_array_equal(np.argmax(scheme, axis=0), bins)


def test_reference_error_behaviour(golden_dir):
    with open(os.path.join(golden_dir, "ref_graph_meta.json")) as f:
        meta = json.load(f)
    assert meta["errors"] == {"one": "IndexError", "nan": "ValueError"}
    with pytest.raises(IndexError):
        O.group_scheme([[np.float32(1.0)]], 10, 1)
    with pytest.raises(ValueError):
        O.group_scheme([[np.float32("nan")]], 10, 1)


@pytest.mark.parametrize("case", ["kat2", "rand_v6", "rand_v12", "tie_v12", "rand_v20"])
def test_reference_graph_vectors(golden_dir, c_oracle, case):
    """view_pooling + group_fusion of the reference (graph code run unmodified) ==
    NumPy oracle == C oracle, bit for bit."""
    z = np.load(os.path.join(golden_dir, "ref_graph_pool_fuse.npz"))
    F, scheme, w, S, P = (z["%s__%s" % (case, k)] for k in ("F", "scheme", "w", "S", "P"))
    V = F.shape[0]
    G = scheme.shape[0]
    np.testing.assert_array_equal(O.group_weight(scheme), w)
    desc = O.view_pooling([F[v] for v in range(V)], scheme)
    for g in range(G):
        np.testing.assert_array_equal(desc[g], P[g])
    np.testing.assert_array_equal(O.group_fusion(desc, w), S)
    bins = np.argmax(scheme, axis=0).astype(np.int32)
    N = F.shape[1]
    Fb = F.reshape(V, N, -1)
    Sc = c_oracle.pool_fuse_fwd(Fb, bins, G, layout="vbd")
    np.testing.assert_array_equal(Sc, S.reshape(N, -1))
    Sb = O.pool_fuse_fwd(np.ascontiguousarray(Fb.transpose(1, 0, 2)), bins, G)
    np.testing.assert_array_equal(Sb, S.reshape(N, -1))


@pytest.mark.parametrize("pool,fill", [("max", 1.0), ("mean", 0.0), ("max", 0.0), ("mean", 1.0)])
@pytest.mark.parametrize("V,G,D", [(6, 10, 24), (12, 8, 40), (20, 16, 8), (80, 4, 12), (1, 1, 5), (5, 2, 1)])
def test_two_oracles_agree(c_oracle, pool, fill, V, G, D):
    rng = np.random.default_rng(V * 100 + G)
    B = 9
    F = rng.standard_normal((B, V, D)).astype(np.float32)
    F[:, :, ::3] = np.maximum(np.round(F[:, :, ::3] * 2) / 2, 0)       # ties
    bins = rng.integers(0, G, (B, V)).astype(np.int32)
    np.testing.assert_array_equal(O.pool_fuse_fwd(F, bins, G, pool, fill),
                                  c_oracle.pool_fuse_fwd(F, bins, G, pool, fill))
    dS = rng.standard_normal((B, D)).astype(np.float32)
    np.testing.assert_array_equal(O.pool_fuse_bwd(dS, F, bins, G, pool),
                                  c_oracle.pool_fuse_bwd(dS, F, bins, G, pool))
    if pool == "max":
        _, m = c_oracle.pool_fuse_fwd(F, bins, G, pool, fill, want_mask=True)
        np.testing.assert_array_equal(m, O.tie_mask_planes(F, bins, G))
    # one scheme shared by the batch (literal score_reduce='batch')
    np.testing.assert_array_equal(O.pool_fuse_fwd(F, bins[0], G, pool, fill),
                                  c_oracle.pool_fuse_fwd(F, bins[0], G, pool, fill))


@pytest.mark.parametrize("pool", ["max", "mean"])
def test_backward_matches_autograd_of_the_literal_graph(pool):
    """TF-autodiff formulas == torch autograd over the op-for-op graph
    (torch.amax shares the gradient among ties like TF's _MinOrMaxGrad)."""
    rng = np.random.default_rng(7)
    B, V, G, D = 5, 12, 8, 16
    F = np.maximum(np.round(rng.standard_normal((B, V, D)) * 2) / 2, 0).astype(np.float32)
    bins = rng.integers(0, G, (B, V)).astype(np.int32)
    dS = rng.standard_normal((B, D)).astype(np.float32)
    Ft = torch.tensor(F, requires_grad=True)
    S = OT.pool_fuse(Ft, bins, G, pool=pool, empty_fill=1.0)
    np.testing.assert_allclose(S.detach().numpy(), O.pool_fuse_fwd(F, bins, G, pool, 1.0), rtol=1e-6, atol=1e-6)
    S.backward(torch.tensor(dS))
    np.testing.assert_allclose(Ft.grad.numpy(), O.pool_fuse_bwd(dS, F, bins, G, pool), rtol=1e-5, atol=1e-7)


def test_score_formula_and_edges(c_oracle):
    """sigmoid(log|x|) == |x|/(1+|x|); exact edge cases of SURVEY H1."""
    x = np.array([0.0, 1.0, -1.0, 2.0 ** 24, -(2.0 ** 25), np.inf, 1e7, 1.0 / 3.0, 3.0, 9.0], dtype=np.float32)
    s_lit = O.score_from_x(x)
    s_rat = O.score_from_x_rational(x)
    np.testing.assert_allclose(s_lit, s_rat, rtol=3e-7, atol=0)
    np.testing.assert_array_equal(c_oracle.score_f32(x), s_rat)
    assert s_rat[0] == 0.0 and s_rat[1] == 0.5 and s_rat[2] == 0.5
    assert s_rat[3] == 1.0 and s_rat[4] == 1.0 and s_rat[5] == 1.0 and s_rat[6] < 1.0
    assert np.isnan(O.score_from_x_rational(np.array([np.nan], dtype=np.float32)))[0]
    for G in (2, 4, 8, 10, 16):
        b = O.bins_from_scores(s_rat, G)
        assert b[0] == 0 and b[1] == G // 2 and b[3] == G          # s == 1.0 -> reference IndexError
        np.testing.assert_array_equal(c_oracle.bins(s_rat, G), b)
    assert O.bins_from_scores(np.array([np.nan], dtype=np.float32), 10)[0] == np.iinfo(np.int32).min
    # x = k/(G-k) puts the true score exactly on edge k/G: flagged as near-edge
    G = 10
    xk = np.array([k / (G - k) for k in range(1, G)], dtype=np.float32)
    assert O.edge_ulps_distance(O.score_from_x_rational(xk), G, k=1).all()


def test_bin_sensitivity_formula_choice():
    """The two float32 formulas differ in the last bit for many inputs yet give the
    same bins away from flagged edges."""
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(200000) * 1.4).astype(np.float32)
    a, b = O.score_from_x(x), O.score_from_x_rational(x)
    for G in (8, 10, 16):
        ba, bb = O.bins_from_scores(a, G), O.bins_from_scores(b, G)
        diff = ba != bb
        assert (~diff | O.edge_ulps_distance(b, G, k=2)).all()
        assert diff.mean() < 1e-4


def test_kernel_order_score_close_to_f64(c_oracle):
    rng = np.random.default_rng(5)
    B, V, Cr = 16, 12, 1024
    R = rng.standard_normal((B, V, Cr)).astype(np.float32)
    lim = np.sqrt(6.0 / (Cr + 1))
    W = rng.uniform(-lim, lim, (V, Cr)).astype(np.float32)
    b = rng.uniform(-1, 1, V).astype(np.float32)
    x64 = c_oracle.view_score_x_f64(R, W, b)
    x64np, _ = O.view_scores(R, W, b, "shape", np.float64)
    np.testing.assert_allclose(x64, x64np, rtol=1e-12, atol=1e-12)
    for E in (1, 4, 8):
        xk = c_oracle.view_score_x_kernel_order(R, W, b, E=E)
        np.testing.assert_allclose(xk, x64, rtol=0, atol=2e-5)


def test_bf16_rounding_helper():
    x = np.array([1.0, 1.00390625, 1.005859375, 3.14159, -2.71828, 65504.0, 1e-40], dtype=np.float32)
    want = torch.tensor(x).to(torch.bfloat16).to(torch.float32).numpy()
    np.testing.assert_array_equal(O.round_bf16(x), want)


def test_graph_literal_cpu_baseline_matches():
    """The torch-CPU graph-literal step used as cpu_baseline computes the same S."""
    rng = np.random.default_rng(11)
    N, V, G, D, Cr = 8, 12, 8, 64, 32
    F = rng.standard_normal((V, N, D)).astype(np.float32)
    R = rng.standard_normal((N, V, Cr)).astype(np.float32)
    W = rng.uniform(-0.3, 0.3, (V, Cr)).astype(np.float32)
    b = rng.uniform(-4, 4, V).astype(np.float32)
    S = OT.reference_step_cpu([torch.tensor(F[v]) for v in range(V)], torch.tensor(R), torch.tensor(W),
                              torch.tensor(b), G).numpy()
    out = O.grouping_fusion_fwd(R, W, b, np.ascontiguousarray(F.transpose(1, 0, 2)), G, score_reduce="batch",
                                score_dtype=np.float32)
    np.testing.assert_allclose(S, out["S"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("case", ["head_v6", "head_v12"])
def test_reference_gvcnn_head_end_to_end(golden_dir, case):
    """The reference's own gvcnn() (nets/model.py:105-166) driven like train.py:264-288 (scores -> its
    group_scheme / group_weight -> shape descriptor, logits) over a stub backbone and NumPy Keras layers:
    the oracle reproduces its scores (float32 rounding of the dot product aside), its scheme and weights
    exactly, and its shape descriptor bit for bit."""
    z = np.load(os.path.join(golden_dir, "ref_graph_head.npz"))
    g = {k: z["%s__%s" % (case, k)] for k in ("R", "W", "b", "F", "scores", "scheme", "weight", "shape_descriptor",
                                               "logits", "cls_w", "cls_b")}
    G, V = g["scheme"].shape
    x, s = O.view_scores(g["R"], g["W"], g["b"], score_reduce="batch", dtype=np.float64)
    np.testing.assert_allclose(s[0], g["scores"], rtol=2e-6, atol=2e-7)   # batch mean cancels: absolute float32 error
    np.testing.assert_array_equal(O.bins_from_scores(s[0].astype(np.float32), G), np.argmax(g["scheme"], axis=0))
    scheme = O.group_scheme([list(g["scores"])], G, V)
    np.testing.assert_array_equal(scheme, g["scheme"])
    np.testing.assert_array_equal(O.group_weight(scheme), g["weight"])
    S = O.group_fusion(O.view_pooling([g["F"][v] for v in range(V)], scheme), O.group_weight(scheme))
    np.testing.assert_array_equal(S, g["shape_descriptor"])
    logits = S.mean(axis=(1, 2), dtype=np.float32) @ g["cls_w"] + g["cls_b"]
    np.testing.assert_allclose(logits, g["logits"], rtol=1e-5, atol=1e-6)
    # basic() (nets/model.py:169-206): plain max over all views = one group, unit weight
    basic = z["%s__basic_descriptor" % case]
    one = O.group_fusion(O.view_pooling([g["F"][v] for v in range(V)], np.ones((1, V), dtype=np.int64)), np.ones(1))
    np.testing.assert_array_equal(one, basic)


def test_power_of_two_mean_shortcut_is_exact_in_float32():
    """The kernels' group mean multiplies by 2^-k instead of dividing when the group size is 2^k
    (csrc/common.cuh mean_of_sum; the backward's reciprocal table): both are the correctly rounded value of the
    same real number, for every float32 - normal, subnormal (also when only the quotient is subnormal), zero,
    infinite.  Checked here on the host for 4 million random bit patterns plus the boundary values per k."""
    rng = np.random.default_rng(5)
    bits = rng.integers(0, 2 ** 32, 4_000_000, dtype=np.uint64).astype(np.uint32)
    edge = np.array([0x00000000, 0x80000000, 0x00000001, 0x00000002, 0x00000003, 0x007fffff, 0x00800000, 0x00800001,
                     0x00ffffff, 0x01000000, 0x7f7fffff, 0x7f800000, 0xff800000, 0x3f800000, 0x3fffffff],
                    dtype=np.uint32)
    x = np.concatenate([bits, edge]).view(np.float32)
    ok = ~np.isnan(x)
    x = x[ok]
    with np.errstate(under="ignore", over="ignore"):
        for k in range(1, 6):                                       # group sizes 2, 4, 8, 16, 32
            n = np.float32(2 ** k)
            r = np.float32(1.0) / n
            assert r * n == 1.0                                     # the reciprocal is exact
            a, b = (x * r).view(np.uint32), (x / n).view(np.uint32)
            assert np.array_equal(a, b), k


def test_add_n_association_orders():
    """The oracle's add_n in both association orders: identical up to 9 terms, (t0 + t1) + (t2 + ... + t9) for the
    reference's num_group = 10, two blocks of 8 for 16; the two orders agree to float32 rounding."""
    rng = np.random.default_rng(0)
    for n in (2, 3, 8, 9):
        t = [rng.standard_normal(50).astype(np.float32) for _ in range(n)]
        np.testing.assert_array_equal(O.add_n(t, "left"), O.add_n(t, "tf8"))
    t = [rng.standard_normal(4000).astype(np.float32) for _ in range(10)]
    blk = t[2]
    for a in t[3:]:
        blk = blk + a
    np.testing.assert_array_equal(O.add_n(t, "tf8"), (t[0] + t[1]) + blk)
    assert np.any(O.add_n(t, "tf8") != O.add_n(t, "left"))
    np.testing.assert_allclose(O.add_n(t, "tf8"), O.add_n(t, "left"), rtol=0, atol=4e-6)
    t = [rng.standard_normal(100).astype(np.float32) for _ in range(16)]
    a = t[0]
    for x in t[1:8]:
        a = a + x
    b = t[8]
    for x in t[9:]:
        b = b + x
    np.testing.assert_array_equal(O.add_n(t, "tf8"), a + b)
