"""Device-side ``log1p`` (di_upload_counts) and the fused tail of predict (di_impute) against the oracle, which is itself
pinned to vectors minted from the reference's own ``MultiNet.predict`` (tests/golden/predict_tail.npz).

Tolerances: the normalised matrix is ``float32(log1p(float64))`` on both sides -- CUDA's and glibc's float64 ``log1p`` may
differ in the last float64 bit, which can flip the float32 rounding in rare cases, so <= 1 float32 ulp is allowed and
the count of such flips is bounded; imputed values are float64 ``expm1`` of float32 inputs: relative 4e-16 (2 ulp)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, synthetic_counts
from oracle.postprocess_oracle import impute_tail, log1p_norm

pytestmark = pytest.mark.gpu
F64_RTOL = 4.5e-16


def _engine(n_pred, H=16, O=12, **kw):
    from deepimpute_b200.engine import Engine
    return Engine(n_pred, hidden=H, sub_outputdim=O, batch_size=32, seed=3, device=0, **kw)


def _partition(rng, G, n_pred, O, dup=0):
    S = len(n_pred)
    uniq = rng.choice(G, S * O - dup, replace=False)
    slots = np.concatenate([uniq, uniq[:dup]]) if dup else uniq
    slots = slots[rng.permutation(len(slots))].astype(np.int32)
    pred_idx = [rng.choice(G, p, replace=False).astype(np.int32) for p in n_pred]
    return pred_idx, slots.reshape(S, O)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_device_log1p_matches_numpy(dtype):
    rng = np.random.default_rng(0)
    N, G = 333, 257                                     # odd sizes: the vector loop and its scalar tail
    raw = rng.poisson(rng.gamma(0.4, 30.0, size=(1, G)), size=(N, G)).astype(dtype)
    raw[5, 7] = 1048575.0
    raw[6, :64] = np.arange(64)
    eng = _engine([20, 9])
    pred_idx, targ = _partition(rng, G, [20, 9], 12)
    eng.set_counts(raw, pred_idx, targ)
    got, want = eng.read_norm(), log1p_norm(raw)
    ulp = np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
    assert ulp.max() <= 1 and (ulp > 0).mean() < 1e-4, (ulp.max(), (ulp > 0).sum())
    # the forward pass sees the same matrix either way
    a = eng.predict()
    eng.set_data(got, pred_idx, targ)
    np.testing.assert_array_equal(a, eng.predict())
    with pytest.raises(RuntimeError):
        eng.impute()                                    # set_data dropped the counts: no silent stale state
    eng.close()


@pytest.mark.parametrize("raw_dtype", [np.float32, np.float64])
@pytest.mark.parametrize("policy", ["restore", "max", None])
def test_impute_reproduces_the_reference_vectors(raw_dtype, policy):
    """External prediction matrix = the fixed matrix the golden vectors were minted with (NaN, overflow, triplicates)."""
    import torch
    z = np.load(os.path.join(GOLDEN, "predict_tail.npz"))
    raw, slots = z["raw"].astype(raw_dtype), z["slot_gene"]
    eng = _engine([6, 6])
    rng = np.random.default_rng(1)
    pred_idx, targ = _partition(rng, raw.shape[1], [6, 6], 12)
    eng.set_counts(raw, pred_idx, targ)
    pred = torch.from_numpy(z["predicted"]).cuda()
    want = z["none" if policy is None else policy]
    got = eng.impute(policy=policy, pred=pred, slot_gene=slots)
    assert got.dtype == np.float64
    np.testing.assert_allclose(got, want, rtol=F64_RTOL, atol=0)
    got32 = eng.impute(policy=policy, pred=pred, slot_gene=slots, dtype=np.float32)
    np.testing.assert_allclose(got32, want.astype(np.float32), rtol=1.2e-7, atol=0)
    # a wider device matrix (leading dimension > n_slots) and ignored (-1) slots
    wide = torch.full((raw.shape[0], len(slots) + 5), 7.0, device="cuda")
    wide[:, :len(slots)] = pred
    s2 = slots.copy(); s2[0] = -1
    got2 = eng.impute(policy=policy, pred=wide, slot_gene=s2)
    want2 = impute_tail(raw, z["predicted"][:, 1:], slots[1:], policy)
    np.testing.assert_allclose(got2, want2, rtol=F64_RTOL, atol=0)
    eng.close()


@pytest.mark.parametrize("math_mode", ["fp32", "tf32x3"])
def test_impute_with_own_forward_many_chunks(math_mode):
    """The engine predicts and imputes chunk by chunk (N > one inference chunk, partial last chunk), pageable output."""
    rng = np.random.default_rng(2)
    N, G, n_pred, O = 16384 + 1000 + 37, 160, [33, 20, 41], 32
    lam = rng.gamma(0.5, 3.0, size=(1, G))
    raw = rng.poisson(lam, size=(N, G)).astype(np.float32)
    eng = _engine(n_pred, H=24, O=O, math_mode=math_mode)
    pred_idx, targ = _partition(rng, G, n_pred, O, dup=7)
    eng.set_counts(raw, pred_idx, targ)
    predicted = eng.predict()
    for policy in ("restore", "max", "none"):
        got = eng.impute(policy=policy)
        want = impute_tail(raw, predicted, targ.reshape(-1), None if policy == "none" else policy)
        np.testing.assert_allclose(got, want, rtol=F64_RTOL, atol=0)
    assert eng.kernel_launches("impute") == 0           # profiling off: nothing recorded, but launches are counted
    n0 = eng.launch_count()
    eng.impute()
    assert eng.launch_count() - n0 >= 2 * 3             # >= 2 chunks x (gather, forward..., impute)
    eng.close()


def test_multinet_fused_route_matches_host_route(test_counts):
    """MultiNet on the reference's example matrix: postprocess='gpu' (counts up, one float64 matrix back) against the
    numpy/pandas route on the same trained engine."""
    from deepimpute_b200.multinet import MultiNet
    raw = test_counts
    net = MultiNet(seed=1234, ncores=1, max_epochs=3, patience=100, verbose=0)
    net.fit(raw)
    assert net.postprocess == "gpu"
    fused = {p: net.predict(raw, policy=p) for p in ("restore", "max", "keep")}
    metrics = dict(net.test_metrics)
    net.postprocess = "host"
    for p, frame in fused.items():
        host = net.predict(raw, policy=p)
        assert list(frame.columns) == list(host.columns) and list(frame.index) == list(host.index)
        np.testing.assert_allclose(frame.values, host.values, rtol=F64_RTOL, atol=0)
    mask = raw.values > 0
    assert np.array_equal(fused["restore"].values[mask], raw.values[mask])
    only = net.predict(raw, imputed_only=True)
    net.postprocess = "gpu"
    only_fused = net.predict(raw, imputed_only=True)
    assert list(only.columns) == list(only_fused.columns)
    # a second fit through the host route reaches the same held-out metrics (same float32 normalised matrix)
    net2 = MultiNet(seed=1234, ncores=1, max_epochs=3, patience=100, verbose=0, postprocess="host")
    net2.fit(raw)
    assert net2.test_metrics["MSE"] == pytest.approx(metrics["MSE"], rel=1e-6)
