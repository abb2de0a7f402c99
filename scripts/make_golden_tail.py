"""Mint tests/golden/predict_tail.npz by running the REFERENCE's own ``MultiNet.predict`` (multinet.py:266-310).

Run in the build container only (needs /root/reference):  python scripts/make_golden_tail.py

What executes is the reference's unmodified ``predict`` body.  Two seams are filled in because the original
dependencies cannot run here:
  * ``self.load()`` (Keras ``model_from_json`` + HDF5, :117-124) returns a stand-in whose ``predict(inputs)`` hands
    back a FIXED prediction matrix -- the network is not what this fixture pins, everything after it is;
  * pandas 3 removed ``DataFrame.groupby(axis=1)`` (:284); the call is routed to the documented equivalent
    ``df.T.groupby(by).mean().T`` for the duration of the run.
Fixture: raw counts [N, G] (float64), targets [S, O] labels with duplicated genes, predicted [N, S*O] float32 containing
NaNs (alone, inside a duplicated gene, in all slots of a gene) and one value above the overflow clamp, and the reference's output for policy restore / max / None.
"""
import os
import sys

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
from make_golden import import_reference, OUT  # noqa: E402


class _Axis1:
    def __init__(self, frame, by):
        self.frame, self.by = frame, np.asarray(by)

    def mean(self):
        return self.frame.T.groupby(by=self.by).mean().T


def reference_predict(ref, raw, targets, predictors, predicted, policy, imputed_only=False):
    """reference MultiNet.predict on ``raw`` with the Keras model replaced by the fixed matrix ``predicted``."""
    net = ref.MultiNet.__new__(ref.MultiNet)
    net.targets, net.predictors = targets, predictors
    O = targets.shape[1]

    class Model:
        def predict(self, inputs):
            assert len(inputs) == len(predictors) and all(x.dtype == np.float32 for x in inputs)
            blocks = [predicted[:, s * O:(s + 1) * O] for s in range(targets.shape[0])]
            return blocks if len(blocks) > 1 else blocks[0]

    net.load = lambda: Model()
    orig = pd.DataFrame.groupby

    def groupby(self, *args, **kwargs):
        if kwargs.get("axis", 0) == 1:
            return _Axis1(self, kwargs.get("by", args[0] if args else None))
        kwargs.pop("axis", None)
        return orig(self, *args, **kwargs)

    pd.DataFrame.groupby = groupby
    try:
        return net.predict(raw, imputed_only=imputed_only, policy=policy)
    finally:
        pd.DataFrame.groupby = orig


def make_case(seed=0, N=24, G=40, S=2, O=12):
    rng = np.random.default_rng(seed)
    raw = rng.poisson(rng.gamma(0.5, 4.0, size=(1, G)), size=(N, G)).astype(np.float64)
    raw[0, 0] = 37.0
    genes = np.array(["g{:02d}".format(j) for j in range(G)], dtype=object)
    cells = np.array(["c{:02d}".format(i) for i in range(N)], dtype=object)
    uniq = rng.choice(G, S * O - 5, replace=False)                 # 5 slots repeat a gene (multinet.py:334-342)
    slots = np.concatenate([uniq, uniq[[0, 0, 1, 2, 3]]])              # one gene three times, three genes twice
    slots = slots[rng.permutation(len(slots))]
    targets = genes[slots].reshape(S, O)
    rest = np.setdiff1d(np.arange(G), uniq)
    predictors = [pd.Index(genes[rng.choice(rest, 6, replace=False)]) for _ in range(S)]
    predicted = rng.gamma(1.0, 0.8, size=(N, S * O)).astype(np.float32)
    predicted[3, 4] = np.nan
    predicted[5, 7] = np.float32(2 * np.log1p(raw.max()) + 1.0)    # above the clamp of :291
    predicted[6, 2] = 0.0
    triple = np.flatnonzero(slots == uniq[0])
    predicted[4, triple[1]] = np.nan                               # NaN inside a duplicated gene: pandas' mean skips it
    predicted[8, np.flatnonzero(slots == uniq[1])] = np.nan        # every slot of a gene NaN -> NaN -> 0 (:291)
    return pd.DataFrame(raw, index=cells, columns=genes), targets, predictors, predicted, slots


def main():
    ref = import_reference()
    raw, targets, predictors, predicted, slots = make_case()
    out = {}
    for name, policy in (("restore", "restore"), ("max", "max"), ("none", None)):
        out[name] = reference_predict(ref, raw, targets, predictors, predicted, policy).values
    only = reference_predict(ref, raw, targets, predictors, predicted, "restore", imputed_only=True)
    np.savez_compressed(os.path.join(OUT, "predict_tail.npz"), raw=raw.values, slot_gene=slots.astype(np.int32),
                        predicted=predicted, restore=out["restore"], max=out["max"], none=out["none"],
                        imputed_only_cols=np.asarray(raw.columns.get_indexer(only.columns), dtype=np.int32),
                        norm32=np.log1p(raw).values.astype(np.float32))
    print("wrote predict_tail.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
