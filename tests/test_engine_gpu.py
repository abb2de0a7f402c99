"""Parity tests proper: the CUDA engine, called through the C-ABI, against the CPU oracle on the same seeded inputs
(same initial weights, same batches, same Philox dropout masks).

Tolerances (stated per north_star "within a stated fp32 tolerance").  Errors are measured against the scale of the
compared tensor, ``max|a - b| / max|b|`` (a relative error per element is meaningless next to a relu threshold or
a cancelling sum):
  fp32 mode  -- CUDA-core FFMA; differs from the oracle only by summation order: 5e-5 of scale for activations and
                predictions (measured: 5e-8 typical; one run out of a dozen on the shared B200 pool showed 2.6e-5 on
                the first forward of the process and was not reproducible, so the bound keeps a margin to it while
                staying 20x below what a single TF32 product would cost), weights after one Adam step within 1e-5
                absolute (lr 1e-3: a wrong-sign update would be 2e-3).
  tf32 mode  -- tcgen05 kind::tf32: the tensor core TRUNCATES fp32 operands to 10 mantissa bits (measured: the
                oracle in operand-truncation mode tracks it to ~2e-4 while the plain fp32 oracle differs by up to
                8e-3), fp32 accumulate.  2e-2 of scale vs the fp32 oracle, 1e-3 vs the truncating oracle.
  tf32x3     -- the default: every GEMM error-compensated (three TF32 products per fp32 product: a_hi b_hi +
                a_hi b_lo + a_lo b_hi), so activations, losses, gradients and predictions sit at 1e-4 of scale or
                better.  Measured end to end (test_multinet_gpu.py): single-pass TF32 gradients drift -- Adam's
                m / sqrt(v) amplifies the relative error of small, cancelling gradient elements -- so all five
                GEMMs are compensated, not only the forward ones.
"""
import numpy as np
import pytest

from deepimpute_b200 import _lib
from deepimpute_b200.engine import Engine, epoch_permutation
from oracle.multinet_oracle import OracleNet, stage

pytestmark = pytest.mark.gpu

FWD_TOL = {"fp32": 5e-5, "tf32": 2e-2, "tf32x3": 1e-4}
MOM_TOL = {"fp32": 2e-5, "tf32": 1e-1, "tf32x3": 2e-4}     # tf32: relu-mask flips leak into dW1/db1
LOSS_TOL = {"fp32": 5e-5, "tf32": 1e-2, "tf32x3": 1e-4}
EPOCH_TOL = {"fp32": 5e-5, "tf32": 2e-2, "tf32x3": 1e-3}       # losses after three epochs of training
PRED_TOL = {"fp32": 1e-4, "tf32": 3e-2, "tf32x3": 2e-3}        # predictions after three epochs of training


def modes():
    lib = _lib.load()
    return [m for m, code in _lib.DI_MATH.items() if lib.di_math_mode_available(code)]


def make_problem(n_cells, n_genes, n_pred, O, seed):
    n_genes = max(n_genes, len(n_pred) * O + max(n_pred) + 16)       # room for disjoint targets and predictors
    rng = np.random.default_rng(seed)
    lam = rng.gamma(0.6, 3.0, size=(1, n_genes)) * rng.gamma(2.0, 0.5, size=(n_cells, 1))
    norm = np.log1p(rng.poisson(lam)).astype(np.float32)
    perm = rng.permutation(n_genes)
    S = len(n_pred)
    targ_idx = perm[:S * O].reshape(S, O).astype(np.int32)
    rest = perm[S * O:] if n_genes > S * O + max(n_pred) else perm
    pred_idx = [rng.choice(rest, p, replace=False).astype(np.int32) for p in n_pred]
    return norm, pred_idx, targ_idx


def pair(n_pred, H, O, B, mode, seed=7, lr=1e-3, rate=0.2, oracle_round=None):
    eng = Engine(n_pred, hidden=H, sub_outputdim=O, learning_rate=lr, batch_size=B, dropout_rate=rate, seed=seed,
                 math_mode=mode)
    ref = OracleNet(n_pred, H, O, learning_rate=lr, batch_size=B, dropout_rate=rate, seed=seed,
                    operand_round=oracle_round)
    return eng, ref


def rel_err(a, b):
    """max|a - b| / max|b|: error against the scale of the tensor."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300)) if b.size else 0.0


SHAPES = [
    # n_pred, H, O, B   (ragged P, H not a multiple of 32 as in reference tests: 150 / 300)
    ([70, 33, 96], 48, 64, 64),
    ([303, 296, 284], 150, 512, 64),           # reference tests/multinet_test.py set-up
    ([1], 8, 8, 4),                            # degenerate: one predictor
    ([600], 256, 512, 64),                     # default topology, one sub-network
    ([129, 2, 64, 31, 257], 300, 96, 32),      # O not a multiple of 32, H = 300 (deepImpute_test.py)
    ([2048, 700], 256, 512, 256),              # BASELINE.json configs[4]: batch 256, predictors capped at 2048
    ([540, 512], 256, 512, 128),               # batch 128: the weight-gradient operands take two passes
]


@pytest.mark.parametrize("mode", modes())
@pytest.mark.parametrize("n_pred,H,O,B", SHAPES)
def test_forward_matches_oracle(mode, n_pred, H, O, B):
    norm, pred_idx, targ_idx = make_problem(200, max(1200, sum(n_pred)), n_pred, O, seed=1)
    eng, ref = pair(n_pred, H, O, B, mode)
    eng.set_data(norm, pred_idx, targ_idx)
    X, _ = stage(norm, pred_idx, targ_idx, np.arange(norm.shape[0]))
    want = np.hstack(ref.forward(X))
    got = eng.predict()
    assert got.shape == want.shape
    assert rel_err(got, want) < FWD_TOL[mode]
    rows = np.array([5, 199, 0, 17, 17], dtype=np.int32)            # arbitrary order, duplicates allowed
    np.testing.assert_allclose(eng.predict(rows=rows), got[rows], rtol=0, atol=0)
    assert eng.predict(rows=np.zeros(0, np.int32)).shape == (0, len(n_pred) * O)
    eng.close()


@pytest.mark.parametrize("mode", modes())
@pytest.mark.parametrize("n_pred,H,O,B", SHAPES)
@pytest.mark.parametrize("nrows", ["full", "partial"])
def test_single_step_matches_oracle(mode, n_pred, H, O, B, nrows):
    """Loss, every intermediate (h, dz2, dz1), Adam first/second moments and the updated weights of one step."""
    n_cells = max(150, B + 22)
    norm, pred_idx, targ_idx = make_problem(n_cells, max(1200, sum(n_pred)), n_pred, O, seed=2)
    eng, ref = pair(n_pred, H, O, B, mode)
    eng.set_data(norm, pred_idx, targ_idx)
    n = B if nrows == "full" else max(1, B - 5)
    rows = np.random.default_rng(3).choice(n_cells, n, replace=False).astype(np.int32)
    X, Y = stage(norm, pred_idx, targ_idx, rows)
    step = 4
    inter = [ref.gradients(s, X[s], Y[s], step)[2] for s in range(len(n_pred))]
    grads = [ref.gradients(s, X[s], Y[s], step)[1] for s in range(len(n_pred))]
    ref.t = step
    loss_ref = ref.train_step(X, Y, step)
    loss = eng.train_step(rows, step)
    assert loss == pytest.approx(loss_ref, rel=LOSS_TOL[mode])

    Hp = -(-H // 32) * 32
    Op = -(-O // 32) * 32
    h, dz2, dz1 = eng.debug_read("h"), eng.debug_read("dz2"), eng.debug_read("dz1")
    for s in range(len(n_pred)):
        hs = h[:n, s * Hp:s * Hp + H]
        # dropout masks are bit-identical: the same units are zero
        assert np.array_equal(hs == 0, inter[s]["h"].numpy() == 0) or mode != "fp32"
        assert rel_err(hs, inter[s]["h"].numpy()) < FWD_TOL[mode]
        assert rel_err(dz2[:n, s * Op:s * Op + O], inter[s]["dz2"].numpy()) < MOM_TOL[mode]
        # dz1 carries the relu mask [z1 > 0]: TF32 noise on a pre-activation next to zero flips the mask of that one
        # unit (its dz1 is then the full dh instead of 0).  Compare where the masks agree; bound the flips.
        d_gpu, d_ref = dz1[:n, s * Hp:s * Hp + H], inter[s]["dz1"].numpy()
        agree = (hs > 0) == (inter[s]["h"].numpy() > 0)
        assert agree.mean() >= {"fp32": 1.0, "tf32": 0.995, "tf32x3": 0.9995}[mode]
        assert rel_err(np.where(agree, d_gpu, 0), np.where(agree, d_ref, 0)) < MOM_TOL[mode]
        # padding rows (beyond the partial batch) and padding columns carry nothing
        assert not dz2[n:, s * Op:(s + 1) * Op].any() and not dz1[n:, s * Hp:(s + 1) * Hp].any()
        assert not dz2[:, s * Op + O:(s + 1) * Op].any() and not h[:, s * Hp + H:(s + 1) * Hp].any()

    for s in range(len(n_pred)):
        bufs, t = eng.get_adam_state(s)
        assert t == step + 1
        for k in range(4):
            g = grads[s][k].numpy().astype(np.float64)
            m_gpu, v_gpu = bufs[2 * k].astype(np.float64), bufs[2 * k + 1].astype(np.float64)
            if not g.any():
                assert not m_gpu.any() and not v_gpu.any()
                continue
            assert rel_err(m_gpu / (1 - float(np.float32(0.9))), g) < MOM_TOL[mode]
            assert rel_err(v_gpu / (1 - float(np.float32(0.999))), g * g) < 2 * MOM_TOL[mode]
    if mode == "fp32":
        for w_gpu, w_ref in zip(eng.get_weights(), ref.get_weights()):
            for a, b in zip(w_gpu, w_ref):
                assert np.max(np.abs(a - b)) < 1e-5            # lr = 1e-3: a wrong-sign update would be 2e-3
    eng.close()


@pytest.mark.parametrize("mode", modes())
def test_epochs_match_oracle(mode):
    """Three epochs with shuffling, a partial last batch and the validation pass: Keras' `loss`/`val_loss`."""
    n_pred, H, O, B = [90, 41], 40, 64, 32
    norm, pred_idx, targ_idx = make_problem(240, 900, n_pred, O, seed=5)
    eng, ref = pair(n_pred, H, O, B, mode, lr=5e-4)
    eng.set_data(norm, pred_idx, targ_idx)
    rng = np.random.default_rng(1)
    cells = rng.permutation(240)
    test_rows, train_rows = cells[:13].astype(np.int32), np.sort(cells[13:]).astype(np.int32)   # 227 = 7*32 + 3
    eng.set_split(train_rows, test_rows)
    Xtr, Ytr = stage(norm, pred_idx, targ_idx, train_rows)
    Xte, Yte = stage(norm, pred_idx, targ_idx, test_rows)
    assert eng.validation_loss() == pytest.approx(ref.loss(Xte, Yte), rel=LOSS_TOL[mode])
    step = 0
    tol = EPOCH_TOL[mode]
    for e in range(3):
        perm = epoch_permutation(7, e, len(train_rows))
        loss_ref, step = ref.train_epoch(Xtr, Ytr, perm, step)
        val_ref = ref.loss(Xte, Yte)
        loss, val = eng.train_epoch(perm)
        assert loss == pytest.approx(loss_ref, rel=tol)
        assert val == pytest.approx(val_ref, rel=tol)
    assert eng.steps_done == step == 24
    want = np.hstack(ref.forward(stage(norm, pred_idx, targ_idx, np.arange(240))[0]))
    assert rel_err(eng.predict(), want) < PRED_TOL[mode]
    eng.close()


@pytest.mark.parametrize("mode", modes())
@pytest.mark.parametrize("B", [128, 256])
def test_epochs_with_large_batches(mode, B):
    """batch 128 / 256 (BASELINE.json configs[4] trains at 256): the epoch graph, N = 256 MMAs, the weight-gradient
    GEMM in several K passes, partial last batch."""
    n_pred, H, O = [96, 40, 130], 40, 64
    n_cells = 3 * B + 57
    norm, pred_idx, targ_idx = make_problem(n_cells, 900, n_pred, O, seed=15)
    eng, ref = pair(n_pred, H, O, B, mode, lr=5e-4)
    eng.set_data(norm, pred_idx, targ_idx)
    cells = np.random.default_rng(2).permutation(n_cells)
    test_rows, train_rows = cells[:20].astype(np.int32), np.sort(cells[20:]).astype(np.int32)
    eng.set_split(train_rows, test_rows)
    Xtr, Ytr = stage(norm, pred_idx, targ_idx, train_rows)
    Xte, Yte = stage(norm, pred_idx, targ_idx, test_rows)
    step = 0
    for e in range(2):
        perm = epoch_permutation(7, e, len(train_rows))
        loss_ref, step = ref.train_epoch(Xtr, Ytr, perm, step)
        loss, val = eng.train_epoch(perm)
        assert loss == pytest.approx(loss_ref, rel=EPOCH_TOL[mode])
        assert val == pytest.approx(ref.loss(Xte, Yte), rel=EPOCH_TOL[mode])
    assert eng.steps_done == step == 8                      # 3 B + 37 training cells: three full batches and a partial one
    want = np.hstack(ref.forward(stage(norm, pred_idx, targ_idx, np.arange(n_cells))[0]))
    assert rel_err(eng.predict(), want) < PRED_TOL[mode]
    eng.close()


@pytest.mark.parametrize("mode", modes())
def test_fit_early_stopping_and_keras_adapters(mode):
    n_pred, H, O, B = [20, 24], 16, 32, 16
    norm, pred_idx, targ_idx = make_problem(120, 400, n_pred, O, seed=8)
    eng, ref = pair(n_pred, H, O, B, mode, lr=2e-3)
    train_rows, test_rows = np.arange(100, dtype=np.int32), np.arange(100, 120, dtype=np.int32)
    Xtr, Ytr = stage(norm, pred_idx, targ_idx, train_rows)
    Xte, Yte = stage(norm, pred_idx, targ_idx, test_rows)
    # Keras-shaped call: lists of arrays, as reference multinet.py:238-244 passes them
    hist = eng.fit_arrays(Xtr, Ytr, (Xte, Yte), epochs=6, patience=2, verbose=0)
    want = ref.fit(Xtr, Ytr, Xte, Yte, epochs=6, patience=2)
    assert len(hist.history["loss"]) == len(want["loss"])
    np.testing.assert_allclose(hist.history["val_loss"], want["val_loss"], rtol=2 * EPOCH_TOL[mode])
    parts = eng.predict_arrays(Xte)                         # list of S arrays [n, O] like model.predict
    assert len(parts) == 2 and parts[0].shape == (20, O)
    assert rel_err(np.hstack(parts), np.hstack(ref.forward(Xte))) < PRED_TOL[mode]
    eng.close()


def test_tf32_tensor_core_rounding_model():
    """kind::tf32 reads fp32 operands and drops the low 13 mantissa bits; the oracle in operand-truncation mode
    must track the tensor-core result much more closely than the plain fp32 oracle does."""
    if "tf32" not in modes():
        pytest.skip("tf32 kernels not built")
    n_pred, H, O, B = [600, 555], 256, 512, 64
    norm, pred_idx, targ_idx = make_problem(256, 2400, n_pred, O, seed=11)
    eng, ref32 = pair(n_pred, H, O, B, "tf32")
    _, ref_tr = pair(n_pred, H, O, B, "fp32", oracle_round="tf32")
    _.close()
    eng.set_data(norm, pred_idx, targ_idx)
    X, _y = stage(norm, pred_idx, targ_idx, np.arange(256))
    got = eng.predict()
    err32 = rel_err(got, np.hstack(ref32.forward(X)))
    err_tr = rel_err(got, np.hstack(ref_tr.forward(X)))
    print("tf32 forward: max rel err vs fp32 oracle {:.2e}, vs truncating oracle {:.2e}".format(err32, err_tr))
    assert err32 < FWD_TOL["tf32"]
    assert err_tr < 1e-3 and err_tr < err32 / 4
    eng.close()
    # the compensated mode removes the truncation error
    eng3, _ref = pair(n_pred, H, O, B, "tf32x3")
    eng3.set_data(norm, pred_idx, targ_idx)
    err3 = rel_err(eng3.predict(), np.hstack(ref32.forward(X)))
    print("tf32x3 forward: max rel err vs fp32 oracle {:.2e}".format(err3))
    assert err3 < FWD_TOL["tf32x3"]
    eng3.close()


def test_odd_batch_size_is_padded_per_batch():
    """batch_size 50 (not a multiple of 32): batches are staged at a padded pitch; results still match the oracle."""
    for mode in modes():
        n_pred, H, O, B = [45, 38], 24, 32, 50
        norm, pred_idx, targ_idx = make_problem(180, 500, n_pred, O, seed=12)
        eng, ref = pair(n_pred, H, O, B, mode, lr=5e-4)
        eng.set_data(norm, pred_idx, targ_idx)
        train_rows, test_rows = np.arange(170, dtype=np.int32), np.arange(170, 180, dtype=np.int32)   # 3*50 + 20
        eng.set_split(train_rows, test_rows)
        Xtr, Ytr = stage(norm, pred_idx, targ_idx, train_rows)
        perm = epoch_permutation(7, 0, 170)
        loss_ref, step = ref.train_epoch(Xtr, Ytr, perm, 0)
        loss, _ = eng.train_epoch(perm)
        assert step == 4 and loss == pytest.approx(loss_ref, rel=EPOCH_TOL[mode])
        eng.close()


def test_sharded_engines_reproduce_the_unsharded_model():
    """Global sub-network ids key init and dropout: two handles owning {0,2} and {1} == one handle owning all."""
    n_pred, H, O, B = [50, 60, 70], 32, 32, 32
    norm, pred_idx, targ_idx = make_problem(100, 500, n_pred, O, seed=9)
    rows = np.arange(32, dtype=np.int32)
    full = Engine(n_pred, hidden=H, sub_outputdim=O, batch_size=B, seed=5, math_mode="fp32")
    full.set_data(norm, pred_idx, targ_idx)
    full.train_step(rows, 0)
    want = full.get_weights()
    for own in ([0, 2], [1]):
        part = Engine([n_pred[s] for s in own], hidden=H, sub_outputdim=O, batch_size=B, seed=5, math_mode="fp32",
                      subnet_ids=own)
        part.set_data(norm, [pred_idx[s] for s in own], targ_idx[own])
        part.train_step(rows, 0)
        for k, s in enumerate(own):
            for a, b in zip(part.get_weights()[k], want[s]):
                np.testing.assert_array_equal(a, b)
        part.close()
    full.close()


def test_predict_device_writes_strided_block():
    import torch
    n_pred, H, O = [40, 30], 16, 32
    norm, pred_idx, targ_idx = make_problem(300, 300, n_pred, O, seed=4)
    eng = Engine(n_pred, hidden=H, sub_outputdim=O, math_mode="fp32")
    eng.set_data(norm, pred_idx, targ_idx)
    want = eng.predict()
    buf = torch.full((300, 100), -1.0, device="cuda")
    eng.predict_device(buf[:, 10:].data_ptr(), 100)
    torch.cuda.synchronize()
    got = buf.cpu().numpy()
    np.testing.assert_array_equal(got[:, 10:10 + 2 * O], want)
    assert (got[:, :10] == -1).all() and (got[:, 10 + 2 * O:] == -1).all()
    eng.close()


def test_errors_are_reported_not_swallowed():
    eng = Engine([10], hidden=8, sub_outputdim=8, math_mode="fp32")
    with pytest.raises(RuntimeError, match="set_data"):
        eng.predict()
    with pytest.raises(RuntimeError, match="no data"):
        eng.train_step(np.arange(4))
    norm = np.ones((20, 30), np.float32)
    with pytest.raises(ValueError):
        eng.set_data(norm, [np.arange(9)], np.arange(8).reshape(1, 8))
    with pytest.raises(ValueError, match="out of range"):
        eng.set_data(norm, [np.arange(10) + 25], np.arange(8).reshape(1, 8))
    eng.set_data(norm, [np.arange(10)], np.arange(10, 18).reshape(1, 8))
    with pytest.raises(RuntimeError, match="1..B"):
        eng.train_step(np.arange(20) % 20, 0) if eng.B < 20 else eng.train_step(np.zeros(0, np.int32), 0)
    with pytest.raises(RuntimeError, match="di_set_split"):
        eng.train_epoch(np.zeros(0, np.int32))
    with pytest.raises(ValueError, match="math_mode"):
        Engine([10], math_mode="fp8")
    eng.close()
    eng.close()                                            # idempotent


def test_save_load_roundtrip(tmp_path):
    n_pred, H, O = [12, 9], 8, 16
    norm, pred_idx, targ_idx = make_problem(64, 200, n_pred, O, seed=6)
    eng = Engine(n_pred, hidden=H, sub_outputdim=O, batch_size=16, math_mode="fp32", seed=3)
    eng.set_data(norm, pred_idx, targ_idx)
    eng.train_step(np.arange(16), 0)
    want = eng.predict()
    path = str(tmp_path / "model.npz")
    eng.save(path, targets=np.array([["a", "b"], ["c", "d"]], dtype=object), predictors=[np.array(["x"]), np.array(["y", "z"])])
    eng.close()
    again, extra = Engine.load(path, math_mode="fp32")
    again.set_data(norm, pred_idx, targ_idx)
    np.testing.assert_array_equal(again.predict(), want)
    assert extra["targets"].tolist() == [["a", "b"], ["c", "d"]]
    assert [list(p) for p in extra["predictors"]] == [["x"], ["y", "z"]]
    again.close()
