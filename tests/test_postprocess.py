"""The element-wise ends of the path on the CPU: the oracle of ``log1p`` / the imputation tail against vectors minted from
the REFERENCE's own ``MultiNet.predict`` (scripts/make_golden_tail.py), against the reference running live when
``/root/reference`` is present, and the host layer's two post-processing routes against each other."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, REFERENCE, ROOT, synthetic_counts
from oracle.postprocess_oracle import impute_tail, log1p_norm

sys.path.insert(0, os.path.join(ROOT, "tests"))
from fake_engine import FakeEngine  # noqa: E402
from deepimpute_b200.multinet import MultiNet  # noqa: E402


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN, "predict_tail.npz"))


@pytest.mark.parametrize("policy", ["restore", "max", "none"])
def test_oracle_tail_reproduces_the_reference_vectors(golden, policy):
    out = impute_tail(golden["raw"], golden["predicted"], golden["slot_gene"], None if policy == "none" else policy)
    assert out.dtype == np.float64
    np.testing.assert_array_equal(out, golden[policy])                  # same numpy, same operations: bit-exact


def test_golden_case_exercises_the_edge_cases(golden):
    slots = golden["slot_gene"]
    assert np.bincount(slots).max() == 3                                 # duplicated target genes (multinet.py:284)
    pred_nan = np.isnan(golden["predicted"])
    assert any(pred_nan[:, slots == g].any(1).sum() and not pred_nan[:, slots == g].all(1).any()
               for g in np.unique(slots) if (slots == g).sum() == 3)     # a NaN the float32 group mean has to skip
    assert np.isnan(golden["predicted"]).any()                           # NaN -> 0 (:291)
    assert (golden["predicted"][np.isfinite(golden["predicted"])] > 2 * np.log1p(golden["raw"].max())).any()
    # the NaN and the overflowing prediction come out as expm1(0) = 0 under policy None unless another slot covers them
    assert np.isfinite(golden["none"]).all()
    np.testing.assert_array_equal(log1p_norm(golden["raw"]), golden["norm32"])


def test_hand_computed_tail():
    raw = np.array([[0.0, 3.0, 0.0], [5.0, 0.0, 1.0]])
    pred = np.array([[1.0, 2.0, 0.5], [0.25, 9.0, np.nan]], dtype=np.float32)       # slots -> genes 0, 0, 2
    out = impute_tail(raw, pred, [0, 0, 2], "restore")
    clamp = 2 * np.log1p(5.0)
    mean1 = np.float32((0.25 + 9.0) / 2)
    assert mean1 > clamp                                                            # row 1, gene 0 is clamped to 0
    expect = np.array([[np.expm1(np.float64(np.float32(1.5))), 3.0, np.expm1(0.5)],
                       [5.0, 0.0, 1.0]])
    np.testing.assert_array_equal(out, expect)
    out_none = impute_tail(raw, pred, [0, 0, 2], None)
    np.testing.assert_allclose(out_none[1], [0.0, np.expm1(np.log1p(0.0)), 0.0], atol=0)
    out_max = impute_tail(raw, pred, [0, 0, 2], "max")
    np.testing.assert_array_equal(out_max[0, 1], 3.0)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="needs /root/reference (build container only)")
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_tail_against_the_live_reference(seed):
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import make_golden_tail as mk
    ref = mk.import_reference()
    raw, targets, predictors, predicted, slots = mk.make_case(seed=seed, N=31, G=57, S=3, O=16)
    for policy in ("restore", "max", None):
        want = mk.reference_predict(ref, raw, targets, predictors, predicted, policy).values
        got = impute_tail(raw.values, predicted, slots, policy)
        np.testing.assert_array_equal(got, want)


class CpuMultiNet(MultiNet):
    def _make_engine(self, inputdims, **kw):
        return FakeEngine(inputdims, **kw)


def test_fused_and_host_routes_of_the_host_layer_agree():
    """``postprocess='gpu'`` drives set_counts / impute, ``'host'`` the numpy restatement; same numbers either way."""
    raw = synthetic_counts(90, 70, seed=4)
    kw = dict(ncores=1, sub_outputdim=16, max_epochs=2, seed=5, verbose=0,
              architecture=[dict(type="dense", neurons=8, activation="relu"), dict(type="dropout", rate=0.2)])
    a = CpuMultiNet(postprocess="gpu", **kw).fit(raw, NN_lim=30, minVMR=0.0)
    b = CpuMultiNet(postprocess="host", **kw).fit(raw, NN_lim=30, minVMR=0.0)
    assert a.test_metrics["MSE"] == pytest.approx(b.test_metrics["MSE"], rel=1e-12)
    for policy in ("restore", "max", "other"):
        np.testing.assert_allclose(a.predict(raw, policy=policy).values, b.predict(raw, policy=policy).values,
                                   rtol=1e-12, atol=0)
    only = a.predict(raw, imputed_only=True)
    assert sorted(only.columns) == sorted(set(a.targets.flatten()))
    with pytest.raises(ValueError):
        MultiNet(ncores=1, postprocess="numpy")
