"""Oracle parity ON THE CONFIGURATIONS THAT ARE BENCHMARKED, in the math mode and launch mode that are benchmarked.

bench.py measures BASELINE.json configs[1] (c2: 10k x 5k, 10 sub-networks) and configs[2] (c3: 50k x 20k, 40
sub-networks) with the default engine: ``tf32x3`` products, one CUDA graph per epoch over 16 sub-network groups,
dependent-launch chain.  The CPU oracle cannot train 40 sub-networks for an epoch in test time, but it does not have
to: the branches of the reference's model share nothing (reference multinet.py:132-148 -- separate Input, Dense,
Dropout, Dense per branch), so a branch trained alone follows exactly the trajectory it follows inside the full
model.  The tests therefore

  1. train the FULL model on the GPU for one epoch of the benchmarked workload (``bench.build_workload``: the very
     matrix, partition and cell split the bench times),
  2. train ``OracleNet(subnet_ids=[...])`` restricted to a few sampled sub-networks on the same rows, same
     permutation, same Philox dropout stream (keyed by the GLOBAL sub-network number) on the CPU,
  3. compare weights, validation loss contribution and predictions of those sub-networks, and
  4. check the independence claim itself on the device: an engine that holds only the sampled sub-networks gives
     bit-identical predictions to the full model's columns, and its loss / val_loss equal the oracle's.

Tolerances (error against the scale of the compared tensor, as in test_engine_gpu.py): predictions 2e-4, weights 5e-4,
losses 1e-4 relative -- the ``tf32x3`` row of DESIGN.md section 4.  Measured on a B200 with short accumulation chains
(kernels_tc.cu, ``acc_sum16``): predictions 1e-6 .. 2e-6, weights 3e-6 .. 5e-5; with one long chain per tile the same
checks sat at 2e-3 .. 4e-3 (profiles/r02e_accumulation.md), which is what these bounds would catch.
"""
import numpy as np
import pytest

from deepimpute_b200.engine import Engine, epoch_permutation
from oracle.multinet_oracle import OracleNet, stage

pytestmark = pytest.mark.gpu

H, O, LR, RATE, SEED = 256, 512, 1e-4, 0.2, 1234


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


def _workload(name):
    import os
    import bench
    os.environ["DI_BENCH_PREDICTORS"] = "0"       # the pinned copy of the raw counts is only for the bench's side blocks
    wl = bench.build_workload(name, "cuda:0")
    norm = wl["norm"].numpy()
    return wl, norm


def _check_sampled(monkeypatch, name, sampled, epochs):
    wl, norm = _workload(name)
    B = wl["B"]
    n_pred = [len(p) for p in wl["pred_idx"]]
    S = len(n_pred)
    assert max(sampled) < S
    tr, te = wl["train_rows"], wl["test_rows"]
    perms = [epoch_permutation(SEED, e, len(tr)) for e in range(epochs)]

    # (1) the full model exactly as bench.py runs it (default math mode, epoch graph, sub-network groups)
    full = Engine(n_pred, hidden=H, sub_outputdim=O, learning_rate=LR, batch_size=B, dropout_rate=RATE, seed=SEED)
    assert full.math_mode == "tf32x3"
    full.set_data(norm, wl["pred_idx"], wl["targ_idx"])
    full.set_split(tr, te)
    full_hist = [full.train_epoch(p) for p in perms]
    assert full.graph_fallbacks() == 0, full.describe()
    rows = np.random.default_rng(5).choice(wl["N"], 1500, replace=False).astype(np.int32)
    full_pred = full.predict(rows=rows)
    full_w = full.get_weights()
    family = full.describe()
    full.close()

    # (4) the sampled sub-networks alone on the device: same trajectory, bit for bit.  The engine picks its kernel
    # family from the size of the optimiser state (and the split-K factor from the number of sub-networks): the small
    # engine is pinned to the family the full model ran with, since families differ in summation order
    monkeypatch.setenv("DEEPIMPUTE_B200_LT", "1" if "fwd/bwd=lt" in family else "0")
    monkeypatch.setenv("DEEPIMPUTE_B200_SPLITK", family.split("splitk=")[1].split()[0])
    sel_pred = [wl["pred_idx"][s] for s in sampled]
    sel_targ = np.ascontiguousarray(wl["targ_idx"][sampled])
    sel_np = [n_pred[s] for s in sampled]
    part = Engine(sel_np, hidden=H, sub_outputdim=O, learning_rate=LR, batch_size=B, dropout_rate=RATE, seed=SEED,
                  subnet_ids=sampled)
    part.set_data(norm, sel_pred, sel_targ)
    part.set_split(tr, te)
    part_hist = [part.train_epoch(p) for p in perms]
    part_pred = part.predict(rows=rows)
    part.close()
    for k, s in enumerate(sampled):
        np.testing.assert_array_equal(part_pred[:, k * O:(k + 1) * O], full_pred[:, s * O:(s + 1) * O])

    # (2) the oracle on the sampled sub-networks
    ref = OracleNet(sel_np, H, O, learning_rate=LR, batch_size=B, dropout_rate=RATE, seed=SEED, subnet_ids=sampled)
    Xtr, Ytr = stage(norm, sel_pred, sel_targ, tr)
    Xte, Yte = stage(norm, sel_pred, sel_targ, te)
    step, ref_hist = 0, []
    for p in perms:
        loss, step = ref.train_epoch(Xtr, Ytr, p, step)
        ref_hist.append((loss, ref.loss(Xte, Yte)))
    del Xtr, Ytr

    # (3) compare
    np.testing.assert_allclose(np.asarray(part_hist), np.asarray(ref_hist), rtol=1e-4)
    want = np.hstack(ref.forward(stage(norm, sel_pred, sel_targ, rows)[0]))
    assert rel_err(part_pred, want) < 2e-4
    for k, s in enumerate(sampled):
        for got_a, ref_a in zip(full_w[s], ref.get_weights()[k]):
            assert rel_err(got_a, ref_a) < 5e-4
    # the full model's own losses are finite, fall, and contain the sampled part
    fh = np.asarray(full_hist)
    assert np.isfinite(fh).all() and (fh[:, 0] > np.asarray(part_hist)[:, 0]).all()
    return fh


def test_c2_sampled_subnetworks_match_oracle(monkeypatch):
    """configs[1]: 10k x 5k, S 10, 149 Adam steps per epoch; two epochs, sub-networks 0, 4 and 9."""
    fh = _check_sampled(monkeypatch, "c2", [0, 4, 9], epochs=2)
    assert fh[1, 0] < fh[0, 0]


def test_c3_sampled_subnetworks_match_oracle(monkeypatch):
    """configs[2], the benchmarked configuration: 50k x 20k, S 40, 743 Adam steps per epoch; one epoch, sub-networks
    3 and 38 (first and last sub-network groups of the epoch graph)."""
    _check_sampled(monkeypatch, "c3", [3, 38], epochs=1)


@pytest.mark.parametrize("knobs", [
    {"DEEPIMPUTE_B200_LT": "0"},                                   # converter-warp kernels (state streams from HBM)
    {"DEEPIMPUTE_B200_LT": "1", "DEEPIMPUTE_B200_SPLITK": "1"},    # every operand by TMA, one CTA per tile
    {"DEEPIMPUTE_B200_LT": "1", "DEEPIMPUTE_B200_SPLITK": "2"},    # ... K loop split over 2 / 4 CTAs
    {"DEEPIMPUTE_B200_LT": "1", "DEEPIMPUTE_B200_SPLITK": "4"},
], ids=["converter", "lt", "lt-split2", "lt-split4"])
def test_kernel_families_follow_the_oracle(monkeypatch, knobs):
    """The engine picks its forward / backward kernel family from the size of the optimiser state (kernels_tc.cu,
    tc_init).  Every family, forced here on one default-topology problem, follows the oracle through an epoch of 40
    Adam steps in graph mode, partial last batch included."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(11)
    n_pred = [540, 513, 600]
    N, G = 40 * 64 - 17 + 128, 2400
    lam = rng.gamma(0.6, 3.0, size=(1, G)) * rng.gamma(2.0, 0.5, size=(N, 1))
    norm = np.log1p(rng.poisson(lam)).astype(np.float32)
    perm = rng.permutation(G)
    targ = perm[:3 * O].reshape(3, O).astype(np.int32)
    pred_idx = [rng.choice(perm[3 * O:], p, replace=False).astype(np.int32) for p in n_pred] if G - 3 * O >= 600 else \
        [rng.choice(G, p, replace=False).astype(np.int32) for p in n_pred]
    tr, te = np.arange(N - 128, dtype=np.int32), np.arange(N - 128, N, dtype=np.int32)
    eng = Engine(n_pred, hidden=H, sub_outputdim=O, learning_rate=1e-3, batch_size=64, dropout_rate=RATE, seed=SEED)
    want_family = "fwd/bwd=lt splitk={}".format(knobs["DEEPIMPUTE_B200_SPLITK"]) if knobs["DEEPIMPUTE_B200_LT"] == "1" else "fwd/bwd=ts"
    assert want_family in eng.describe(), eng.describe()
    eng.set_data(norm, pred_idx, targ)
    eng.set_split(tr, te)
    ref = OracleNet(n_pred, H, O, learning_rate=1e-3, batch_size=64, dropout_rate=RATE, seed=SEED)
    Xtr, Ytr = stage(norm, pred_idx, targ, tr)
    Xte, Yte = stage(norm, pred_idx, targ, te)
    step = 0
    for epoch in range(2):
        p = epoch_permutation(SEED, epoch, len(tr))
        got = eng.train_epoch(p)
        loss, step = ref.train_epoch(Xtr, Ytr, p, step)
        np.testing.assert_allclose(got, (loss, ref.loss(Xte, Yte)), rtol=1e-4)
    assert eng.graph_fallbacks() == 0
    want = np.hstack(ref.forward(stage(norm, pred_idx, targ, np.arange(N))[0]))
    assert rel_err(eng.predict(), want) < 1e-4          # measured 1e-6; a single long chain per tile gives 2.8e-3 here
    for got_w, ref_w in zip(eng.get_weights(), ref.get_weights()):
        for a, b in zip(got_w, ref_w):
            assert rel_err(a, b) < 5e-4                 # measured 3e-5
    eng.close()
