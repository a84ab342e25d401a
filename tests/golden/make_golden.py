#!/usr/bin/env python
"""Generate tests/golden/ref_graph_*.npz by running the REFERENCE'S OWN code.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box); the .npz files it writes are committed and are what the tests
read.  Usage:  python tests/golden/make_golden.py

What is executed:
  * ``group_scheme`` / ``group_weight`` (nets/model.py:16-41) - pure NumPy
    functions of the reference, run unmodified (``np.int`` is aliased to
    ``int`` because NumPy >= 1.24 removed it).
  * ``view_pooling`` / ``group_fusion`` (nets/model.py:44-102) - the
    reference's graph-construction code, run unmodified over ``_FakeTF``: an
    eager NumPy stand-in for exactly the 13 TensorFlow ops those two functions
    call.  TensorFlow itself cannot be installed here, so the op *kernels* are
    ours (documented TF semantics, float32, one rounding per op, add_n left to
    right); the op *graph* - which views are gathered, the ones dummy, the
    order of the weighted sum - is the reference's.  That is the strongest
    pin available: "parity unpinned" w.r.t. TF's Eigen kernels, pinned w.r.t.
    the reference's control flow.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


class _FakeTF(types.ModuleType):
    """Eager NumPy versions of the TF ops nets/model.py:44-102 uses."""

    def __init__(self):
        super().__init__("tensorflow")
        self.contrib = types.SimpleNamespace(slim=types.SimpleNamespace())
        self.compat = types.SimpleNamespace(v1=types.SimpleNamespace(AUTO_REUSE=object()))
        self.math = types.SimpleNamespace(log=np.log)
        self.nn = types.SimpleNamespace(sigmoid=lambda x: 1 / (1 + np.exp(-x)))
        self.keras = types.SimpleNamespace()

    @staticmethod
    def _t(x):                      # convert_to_tensor: a list of tensors packs
        return np.stack([np.asarray(e) for e in x]) if isinstance(x, (list, tuple)) else np.asarray(x)

    def ones_like(self, x):
        return np.ones_like(self._t(x))

    def unstack(self, x):
        return [e for e in self._t(x)]

    def where(self, cond):
        return np.argwhere(self._t(cond))

    def squeeze(self, x, axis=None):
        return np.squeeze(x, axis=axis)

    def size(self, x):
        return np.asarray(x).size

    def greater(self, a, b):
        return a > b

    def cond(self, pred, true_fn, false_fn):
        return true_fn() if bool(pred) else false_fn()

    def gather(self, params, indices):
        return self._t(params)[np.asarray(indices)]

    def reduce_max(self, x, axis=None):
        return self._t(x).max(axis=axis)

    def multiply(self, a, b):
        return np.asarray(a) * np.asarray(b)

    def reduce_sum(self, x):
        acc = np.float32(0)
        for e in self._t(x).reshape(-1):
            acc = np.float32(acc + e)
        return acc

    def add_n(self, xs):
        acc = xs[0]
        for t in xs[1:]:
            acc = acc + t
        return acc

    def div(self, a, b):
        return np.asarray(a) / b

    def abs(self, x):
        return np.abs(x)


def load_reference_model():
    if not hasattr(np, "int"):
        np.int = int                                  # nets/model.py:21
    fake = _FakeTF()
    sys.modules["tensorflow"] = fake
    nets = types.ModuleType("nets")
    nets.__path__ = []
    nets.inception_v3 = types.ModuleType("nets.inception_v3")
    nets.resnet_v2 = types.ModuleType("nets.resnet_v2")
    sys.modules["nets"] = nets
    sys.modules["nets.inception_v3"] = nets.inception_v3
    sys.modules["nets.resnet_v2"] = nets.resnet_v2
    spec = importlib.util.spec_from_file_location("nets.model", os.path.join(REF, "nets", "model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference_model()
    rng = np.random.default_rng(20261017)

    # ---- host part: group_scheme / group_weight, num_group = 10 (train.py:97)
    host = {"scores": [], "scheme": [], "weight": [], "error": []}
    score_sets = [
        [0.05, 0.15, 0.25, 0.35, 0.45, 0.55],                       # spread
        [0.5, 0.5, 0.5, 0.5, 0.5, 0.5],                             # |x| = 1 -> one group
        [0.0, 0.0999999, 0.1, 0.8999999, 0.9, 0.99999994],          # exact edges
        [0.3, 0.3, 0.7, 0.7, 0.7, 0.1, 0.1, 0.1, 0.1, 0.95, 0.2, 0.6],  # V = 12
    ]
    for _ in range(8):
        score_sets.append(list(rng.random(12, dtype=np.float32)))
    for sc in score_sets:
        sc32 = [np.float32(s) for s in sc]
        V = len(sc32)
        scheme = ref.group_scheme([sc32], 10, V)
        host["scores"].append(np.asarray(sc32, dtype=np.float32))
        host["scheme"].append(np.asarray(scheme, dtype=np.int64))
        host["weight"].append(ref.group_weight(scheme))
    # reference error behaviour: score == 1.0 -> IndexError; NaN -> ValueError
    errs = {}
    for name, val in (("one", np.float32(1.0)), ("nan", np.float32("nan"))):
        try:
            ref.group_scheme([[val]], 10, 1)
            errs[name] = "none"
        except Exception as e:                                     # noqa: BLE001
            errs[name] = type(e).__name__
    np.savez(os.path.join(HERE, "ref_graph_host.npz"),
             n=len(score_sets),
             **{"scores_%d" % i: a for i, a in enumerate(host["scores"])},
             **{"scheme_%d" % i: a for i, a in enumerate(host["scheme"])},
             **{"weight_%d" % i: a for i, a in enumerate(host["weight"])})

    # ---- graph part: view_pooling + group_fusion (shipped max / ones variant)
    cases = {}
    # KAT-2: unit_test.py:18-19 data through nets/model.py
    F = np.array([[8, 1, 220, 55], [3, 4, 3, -1], [54, 1, 6, -53], [-3, -4, 35, -1], [0, 34, 0, -23]],
                 dtype=np.float32)
    scheme = np.array([[0, 1, 0, 0, 0], [0, 0, 1, 0, 0], [0, 0, 0, 0, 0], [1, 0, 0, 1, 1], [0, 0, 0, 0, 0]],
                      dtype=np.int32)
    w = ref.group_weight(scheme)
    desc = ref.view_pooling([f[None] for f in F], scheme)          # V views of shape [N=1, 4]
    S = ref.group_fusion(desc, w)
    cases["kat2"] = dict(F=F[:, None, :], scheme=scheme, w=w, S=S,
                         P=np.stack([desc[g] for g in range(5)]))
    # random small batches, G = 10, bins drawn through the reference's own
    # group_scheme from random scores; descriptors are [N, h, w, C] maps.
    for name, (V, N, shape, tie) in {
        "rand_v6": (6, 4, (2, 2, 8), False),
        "rand_v12": (12, 3, (1, 1, 64), False),
        "tie_v12": (12, 3, (1, 1, 64), True),
        "rand_v20": (20, 2, (32,), False),
    }.items():
        sc = [np.float32(s) for s in rng.random(V, dtype=np.float32)]
        scheme = ref.group_scheme([sc], 10, V)
        w = ref.group_weight(scheme)
        views = []
        for v in range(V):
            x = rng.standard_normal((N,) + shape).astype(np.float32)
            if tie:
                x = np.maximum(np.round(x * 2) / 2, 0).astype(np.float32)
            views.append(x)
        desc = ref.view_pooling(views, scheme)
        S = ref.group_fusion(desc, w)
        cases[name] = dict(F=np.stack(views), scheme=np.asarray(scheme, dtype=np.int32), w=w, S=S,
                           P=np.stack([desc[g] for g in range(10)]),
                           scores=np.asarray(sc, dtype=np.float32))
    flat = {}
    for cname, d in cases.items():
        for k, v in d.items():
            flat["%s__%s" % (cname, k)] = v
    np.savez(os.path.join(HERE, "ref_graph_pool_fuse.npz"), **flat)

    # ---- the whole head: the reference's own gvcnn() (nets/model.py:105-166) run end to end, twice, exactly
    #      like train.py:264-288 drives it: first for the view scores, then - after its own group_scheme /
    #      group_weight on the host - for the shape descriptor and the logits.  The backbone is a stub that
    #      hands out pre-generated block3 / block4 maps; the Keras layers are NumPy stand-ins with seeded
    #      weights (Dense(1) per view inside the loop, model.py:145; Dense(num_classes) at :164).
    head_cases = {}
    for name, (V, N, hw3, C3, hw4, C4, ncls) in {"head_v6": (6, 4, 3, 64, 2, 128, 5),
                                                  "head_v12": (12, 3, 2, 96, 1, 256, 7)}.items():
        feats3 = [rng.standard_normal((N, hw3, hw3, C3)).astype(np.float32) for _ in range(V)]
        feats4 = [np.maximum(rng.standard_normal((N, hw4, hw4, C4)), 0).astype(np.float32) for _ in range(V)]
        lim = np.sqrt(6.0 / (C3 + 1))
        dense_w = [rng.uniform(-lim, lim, (C3, 1)).astype(np.float32) for _ in range(V)]
        dense_b = [np.float32(rng.uniform(-3, 3)) for _ in range(V)]         # spread the batch-mean scores
        cls_w = (rng.standard_normal((C4, ncls)) * 0.05).astype(np.float32)
        cls_b = np.zeros(ncls, dtype=np.float32)
        state = {"view": 0, "dense": 0}

        class _Inputs(np.ndarray):                                            # inputs.get_shape().as_list()
            def get_shape(self):
                return types.SimpleNamespace(as_list=lambda: list(self.shape))

        def resnet_v2_50(batch_view, num_classes=None, is_training=None, reuse=None):
            v = state["view"]
            state["view"] += 1
            return None, {"resnet_v2_50/block3": feats3[v], "resnet_v2_50/block4": feats4[v]}

        class _GAP:
            def __call__(self, x):
                return np.asarray(x, dtype=np.float32).mean(axis=(1, 2), dtype=np.float32)

        class _Dense:
            def __init__(self, units):
                i = state["dense"]
                state["dense"] += 1
                self.k, self.b = (dense_w[i], dense_b[i]) if units == 1 else (cls_w, cls_b)

            def __call__(self, x):
                return (np.asarray(x, dtype=np.float32) @ self.k + self.b).astype(np.float32)

        import contextlib
        tfm = sys.modules["tensorflow"]
        tfm.transpose = lambda x, perm: np.transpose(np.asarray(x), perm)
        tfm.reduce_mean = lambda x: np.float32(np.asarray(x, dtype=np.float32).mean(dtype=np.float32))
        tfm.nn = types.SimpleNamespace(sigmoid=lambda x: np.float32(1) / (np.float32(1) + np.exp(-x, dtype=np.float32)))
        tfm.math = types.SimpleNamespace(log=lambda x: np.log(x, dtype=np.float32))
        tfm.keras = types.SimpleNamespace(layers=types.SimpleNamespace(GlobalAveragePooling2D=_GAP, Dense=_Dense))
        ref.slim = types.SimpleNamespace(arg_scope=lambda scope: contextlib.nullcontext())
        ref.resnet_v2 = types.SimpleNamespace(resnet_arg_scope=lambda: None, resnet_v2_50=resnet_v2_50)
        images = np.zeros((N, V, 4, 4, 3), dtype=np.float32).view(_Inputs)   # the stub backbone ignores pixels

        def run(scheme, weight):
            state["view"] = state["dense"] = 0
            return ref.gvcnn(images, ncls, scheme, weight, is_training=False, dropout_keep_prob=1.0)

        G = 10
        scores, _, _ = run(np.zeros((G, V), dtype=np.int32), np.ones(G, dtype=np.float32))    # partial_run #1
        scheme = ref.group_scheme([scores], G, V)                                             # train.py:277
        weight = ref.group_weight(scheme)                                                     # train.py:278
        scores2, shape_descriptor, logits = run(scheme, weight)                               # partial_run #2
        assert all(a == b for a, b in zip(scores, scores2))
        state["view"] = state["dense"] = 0
        basic_descriptor, _ = ref.basic(images, ncls, is_training=False, dropout_keep_prob=1.0)   # model.py:169-206
        head_cases[name] = dict(
            basic_descriptor=basic_descriptor,
            R=np.stack([f.mean(axis=(1, 2), dtype=np.float32) for f in feats3], axis=1),      # [N, V, C3] post-GAP
            W=np.stack([k[:, 0] for k in dense_w]), b=np.asarray(dense_b, dtype=np.float32),
            F=np.stack(feats4), scores=np.asarray(scores, dtype=np.float32),
            scheme=np.asarray(scheme, dtype=np.int32), weight=weight,
            shape_descriptor=shape_descriptor, logits=logits, cls_w=cls_w, cls_b=cls_b)
    flat = {}
    for cname, d in head_cases.items():
        for k, v in d.items():
            flat["%s__%s" % (cname, k)] = v
    np.savez(os.path.join(HERE, "ref_graph_head.npz"), **flat)
    print("head scores:", head_cases["head_v6"]["scores"], "bins:", np.argmax(head_cases["head_v6"]["scheme"], axis=0))

    with open(os.path.join(HERE, "ref_graph_meta.json"), "w") as f:
        json.dump({"generated_by": "tests/golden/make_golden.py",
                   "reference": "ace19-dev/gvcnn-tf nets/model.py (group_scheme, group_weight, "
                                "view_pooling, group_fusion run unmodified over a NumPy stand-in for TF ops)",
                   "errors": errs, "cases": sorted(cases), "head_cases": sorted(head_cases)}, f, indent=1)
    print("errors:", errs)
    print("kat2 S:", cases["kat2"]["S"], "w:", cases["kat2"]["w"])


if __name__ == "__main__":
    main()
