"""BASELINE.json configs[1] at full size (10k cells x 5k genes, 10 sub-networks of the default topology) through
size-independent properties -- the CPU oracle would need minutes here, so the checks are relations the path must keep:

* the two independent CUDA implementations (fp32 CUDA cores, error-compensated TF32 tensor cores) agree on losses and
  predictions after training on the same batches, within the tolerance stated in DESIGN.md section 4;
* training makes progress (epoch losses fall) and the validation loss is the wMSE of the predictions themselves;
* prediction is per-cell: any order / subset of rows gives bit-identical rows;
* the fused tail keeps observed counts (policy "restore"), dominates them (policy "max") and is finite everywhere.
"""
import numpy as np
import pytest

from conftest import synthetic_counts
from deepimpute_b200.engine import Engine, epoch_permutation

pytestmark = pytest.mark.gpu

N, G, S, H, O, B = 10_000, 5_000, 10, 256, 512, 64


@pytest.fixture(scope="module")
def problem():
    raw = synthetic_counts(N, G, seed=0, rank=32).values.astype(np.float32)
    rng = np.random.default_rng(1)
    perm = rng.permutation(G)
    targ = perm[:S * O - 400]
    targ = np.concatenate([targ, targ[:400]]).reshape(S, O).astype(np.int32)      # 4720 genes in 5120 slots: duplicates
    n_pred = [int(p) for p in rng.integers(540, 660, size=S)]
    pred_idx = [rng.choice(G, p, replace=False).astype(np.int32) for p in n_pred]
    cells = rng.permutation(N)
    n_test = int(0.05 * N)                                                        # multinet.py:228
    return dict(raw=raw, targ=targ, n_pred=n_pred, pred_idx=pred_idx,
                test_rows=cells[:n_test].astype(np.int32), train_rows=np.sort(cells[n_test:]).astype(np.int32))


def _train(problem, mode, epochs=2):
    eng = Engine(problem["n_pred"], hidden=H, sub_outputdim=O, batch_size=B, seed=1234, math_mode=mode)
    eng.set_counts(problem["raw"], problem["pred_idx"], problem["targ"])
    eng.set_split(problem["train_rows"], problem["test_rows"])
    hist = [eng.train_epoch(epoch_permutation(1234, e, len(problem["train_rows"]))) for e in range(epochs)]
    return eng, np.asarray(hist)


def test_config2_properties(problem):
    fp32, h32 = _train(problem, "fp32")
    x3, hx3 = _train(problem, "tf32x3")
    steps = -(-len(problem["train_rows"]) // B)
    assert fp32.steps_done == x3.steps_done == 2 * steps == 2 * 149
    # both implementations follow the same trajectory; losses fall
    np.testing.assert_allclose(hx3, h32, rtol=1e-3)
    assert h32[1, 0] < h32[0, 0] and h32[1, 1] < h32[0, 1]
    # predictions of all cells agree to the stated tolerance (error against the scale of the tensor)
    p32, px3 = fp32.predict(), x3.predict()
    assert p32.shape == (N, S * O) and np.isfinite(px3).all() and (px3 >= 0).all()       # softplus output
    assert np.max(np.abs(px3 - p32)) / np.max(np.abs(p32)) < 2e-3
    # val_loss is the weighted MSE of those predictions on the held-out cells (multinet.py:36-41, summed over branches)
    te = problem["test_rows"]
    y = np.log1p(problem["raw"][np.ix_(te, problem["targ"].reshape(-1))].astype(np.float64))
    val = sum(np.mean(y[:, s * O:(s + 1) * O] * (y[:, s * O:(s + 1) * O] - px3[te, s * O:(s + 1) * O]) ** 2) for s in range(S))
    assert x3.validation_loss() == pytest.approx(val, rel=1e-4)
    # per-cell: order and subset of rows do not matter
    rows = np.random.default_rng(3).permutation(N)[:3000].astype(np.int32)
    np.testing.assert_array_equal(x3.predict(rows=rows), px3[rows])
    # fused tail
    raw = problem["raw"]
    restored = x3.impute(policy="restore", dtype=np.float32)
    assert np.isfinite(restored).all() and (restored >= 0).all()
    assert np.array_equal(restored[raw > 0], raw[raw > 0])
    untouched = np.setdiff1d(np.arange(G), problem["targ"].reshape(-1))
    assert np.array_equal(restored[:, untouched], raw[:, untouched])                      # genes nobody imputes
    biggest = x3.impute(policy="max", dtype=np.float32)
    assert (biggest >= raw).all() and (biggest >= restored - 1e-6 * np.abs(restored)).all()
    fp32.close(); x3.close()
