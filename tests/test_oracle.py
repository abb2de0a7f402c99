"""Guards on the CPU oracle itself (oracle/): the neural-network arithmetic is "parity unpinned" by the reference
(its tests hold no golden vector for it), so the oracle is pinned against (1) the published
Philox4x32-10 known-answer vectors, (2) hand-derived fp64 cases, (3) torch autograd as an independent
differentiator, (4) finite differences, (5) the Keras/TF Adam and EarlyStopping rules written out long-hand."""
import numpy as np
import pytest
import torch

from oracle.multinet_oracle import OracleNet, epoch_permutation, glorot_weights, stage
from oracle.philox import dropout_keep_mask, dropout_threshold, philox4x32_10


# ---- Philox: known-answer vectors published with Random123 (kat_vectors, philox4x32 10 rounds) ------------------
@pytest.mark.parametrize("ctr,key,expect", [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
])
def test_philox_known_answers(ctr, key, expect):
    out = philox4x32_10(*[np.uint32(c) for c in ctr], *key)
    assert tuple(int(np.asarray(o).reshape(-1)[0]) for o in out) == expect


def test_dropout_mask_definition():
    rate, seed = 0.2, (7 << 32) | 99
    m = dropout_keep_mask(seed, 3, 11, 10, 6, rate)
    thr = dropout_threshold(rate)
    for b in range(10):
        for j in range(6):
            w = philox4x32_10(np.uint32(j), np.uint32(b >> 2), np.uint32(3), np.uint32(11), 99, 7)
            assert m[b, j] == (int(w[b & 3]) >= thr)
    big = dropout_keep_mask(seed, 0, 0, 512, 256, rate)
    assert abs(big.mean() - 0.8) < 0.01
    assert dropout_keep_mask(seed, 0, 0, 4, 4, 0.0).all()
    # different sub-network / step -> different mask
    assert (dropout_keep_mask(seed, 0, 0, 64, 64, rate) != dropout_keep_mask(seed, 1, 0, 64, 64, rate)).any()
    assert (dropout_keep_mask(seed, 0, 0, 64, 64, rate) != dropout_keep_mask(seed, 0, 1, 64, 64, rate)).any()


# ---- forward / loss: hand-derived fp64 case (P 3, H 2, O 2, B 2) -------------------------------------------------
def _tiny():
    net = OracleNet([3], 2, 2, learning_rate=0.1, batch_size=2, dropout_rate=0.0, seed=1, dtype=torch.float64)
    W1 = np.array([[1.0, -1.0], [0.5, 0.25], [-2.0, 1.0]])
    b1 = np.array([0.1, -0.2])
    W2 = np.array([[1.0, -0.5], [0.25, 2.0]])
    b2 = np.array([0.0, 0.3])
    net.set_weights([(W1, b1, W2, b2)])
    x = np.array([[1.0, 2.0, 0.0], [0.0, 1.0, 1.0]])
    y = np.array([[0.0, 1.5], [2.0, 0.5]])
    return net, x, y


def test_forward_and_wmse_by_hand():
    net, x, y = _tiny()
    # z1 = x W1 + b1 = [[2.1, -0.7], [-1.4, 1.05]] -> relu [[2.1, 0], [0, 1.05]]
    # z2 = h W2 + b2 = [[2.1, -0.75], [0.2625, 2.4]]
    z2 = np.array([[2.1, -1.05 + 0.3], [0.2625, 2.1 + 0.3]])
    yhat = np.log1p(np.exp(z2))
    out = net.forward([x])[0]
    np.testing.assert_allclose(out, yhat, rtol=1e-14)
    # wMSE (reference multinet.py:36-41): mean over all 4 entries of y (y - yhat)^2
    expect = np.mean(y * (y - yhat) ** 2)
    assert net.loss([x], [y]) == pytest.approx(expect, rel=1e-14)
    # zeros in y carry zero weight: changing the prediction where y == 0 cannot change the loss
    L, grads, inter = net.gradients(0, x, y, 0, training=False)
    assert inter["dz2"][0, 0] == 0.0


def test_gradients_match_autograd_and_finite_differences():
    rng = np.random.default_rng(3)
    P, H, O, B = 7, 5, 4, 6
    net = OracleNet([P], H, O, learning_rate=1e-3, batch_size=B, dropout_rate=0.3, seed=5, dtype=torch.float64)
    x = rng.gamma(1.0, 1.0, size=(B, P))
    y = np.log1p(rng.poisson(2.0, size=(B, O))).astype(np.float64)
    L, grads, inter = net.gradients(0, x, y, step=9, training=True)

    # independent differentiator: torch autograd over the same forward with the same (Philox) mask
    keep = torch.from_numpy(dropout_keep_mask(5, 0, 9, B, H, 0.3))
    ws = [w.clone().requires_grad_(True) for w in net.w[0]]
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    a = torch.relu(xt @ ws[0] + ws[1])
    h = torch.where(keep, a / (1 - 0.3), torch.zeros_like(a))
    yhat = torch.nn.functional.softplus(h @ ws[2] + ws[3])
    loss = (yt * (yt - yhat) ** 2).mean()
    loss.backward()
    assert L == pytest.approx(float(loss), rel=1e-13)
    for g, w in zip(grads, ws):
        np.testing.assert_allclose(g.numpy(), w.grad.numpy(), rtol=1e-10, atol=1e-14)

    # finite differences on a few entries of each tensor
    def loss_at(k, idx, eps):
        old = net.w[0][k][idx].item()
        net.w[0][k][idx] = old + eps
        val = net.gradients(0, x, y, step=9, training=True)[0]
        net.w[0][k][idx] = old
        return val
    for k, idx in [(0, (2, 1)), (0, (6, 4)), (1, (3,)), (2, (4, 2)), (3, (1,))]:
        fd = (loss_at(k, idx, 1e-6) - loss_at(k, idx, -1e-6)) / 2e-6
        assert grads[k][idx].item() == pytest.approx(fd, rel=1e-5, abs=1e-9)


# ---- Adam: TensorFlow ResourceApplyAdam rule written out --------------------------------------------------------
def test_adam_rule_and_shared_step_counter():
    rng = np.random.default_rng(0)
    n_pred, H, O, B = [4, 6], 3, 5, 8
    net = OracleNet(n_pred, H, O, learning_rate=1e-2, batch_size=B, dropout_rate=0.0, seed=2, dtype=torch.float64)
    w0 = net.get_weights()
    X = [rng.gamma(1.0, 1.0, size=(B, p)) for p in n_pred]
    Y = [np.log1p(rng.poisson(3.0, size=(B, O))).astype(np.float64) for _ in n_pred]
    m = [[np.zeros_like(a) for a in ws] for ws in w0]
    v = [[np.zeros_like(a) for a in ws] for ws in w0]
    w = [[a.astype(np.float64).copy() for a in ws] for ws in w0]
    b1, b2, eps, lr = float(np.float32(0.9)), float(np.float32(0.999)), float(np.float32(1e-7)), 1e-2
    for t in (1, 2, 3):
        grads = [net.gradients(s, X[s], Y[s], t - 1, training=True)[1] for s in range(2)]
        total = net.train_step(X, Y, t - 1)
        assert total == pytest.approx(sum(net_loss for net_loss in
                                          [float(np.mean(Y[s] * (Y[s] - _fwd(w[s], X[s])) ** 2)) for s in range(2)]),
                                      rel=1e-12)
        lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        for s in range(2):
            for k in range(4):
                g = grads[s][k].numpy()
                m[s][k] = b1 * m[s][k] + (1 - b1) * g
                v[s][k] = b2 * v[s][k] + (1 - b2) * g * g
                w[s][k] = w[s][k] - lr_t * m[s][k] / (np.sqrt(v[s][k]) + eps)
                np.testing.assert_allclose(net.w[s][k].numpy(), w[s][k], rtol=1e-9, atol=1e-12)
    assert net.t == 3
    # first step moves every weight with a non-negligible gradient by ~lr (Adam's signature)
    d = np.abs(net.get_weights()[0][2] - w0[0][2])
    assert np.all(d < 3.5 * lr)


def _fwd(ws, x):
    h = np.maximum(x @ ws[0] + ws[1], 0)
    return np.log1p(np.exp(h @ ws[2] + ws[3]))


# ---- epoch loop: Keras fit() bookkeeping -------------------------------------------------------------------------
def test_epoch_loss_is_sample_weighted_and_partial_batch_kept():
    rng = np.random.default_rng(1)
    n, P, H, O, B = 23, 5, 4, 3, 8          # batches of 8, 8, 7
    X = [rng.gamma(1.0, 1.0, size=(n, P)).astype(np.float32)]
    Y = [np.log1p(rng.poisson(2.0, size=(n, O))).astype(np.float32)]
    a = OracleNet([P], H, O, batch_size=B, dropout_rate=0.2, seed=4)
    b = OracleNet([P], H, O, batch_size=B, dropout_rate=0.2, seed=4)
    perm = epoch_permutation(4, 0, n)
    assert sorted(perm.tolist()) == list(range(n))
    loss, step = a.train_epoch(X, Y, perm, 0)
    assert step == 3 and a.t == 3
    acc = 0.0
    for i, lo in enumerate(range(0, n, B)):
        rows = perm[lo:lo + B]
        acc += b.train_step([X[0][rows]], [Y[0][rows]], i) * len(rows)
    assert loss == pytest.approx(acc / n, rel=1e-12)
    for wa, wb in zip(a.get_weights()[0], b.get_weights()[0]):
        np.testing.assert_array_equal(wa, wb)


def test_early_stopping_rule(monkeypatch):
    net = OracleNet([3], 2, 2, batch_size=4, dropout_rate=0.0, seed=0)
    vals = iter([5.0, 4.0, 4.5, 4.0, 4.2, 4.1, 4.3, 3.0, 2.0])
    monkeypatch.setattr(net, "train_epoch", lambda X, Y, perm, step: (1.0, step + 1))
    monkeypatch.setattr(net, "loss", lambda X, Y: next(vals))
    hist = net.fit([np.zeros((4, 3), np.float32)], [np.zeros((4, 2), np.float32)], [], [], epochs=50, patience=5)
    # best = 4.0 at epoch 2; epochs 3..7 do not beat it (4.0 is not < 4.0) -> stop after 7 epochs
    assert len(hist["val_loss"]) == 7
    vals = iter([5.0, 4.0, 3.0])
    hist = net.fit([np.zeros((4, 3), np.float32)], [np.zeros((4, 2), np.float32)], [], [], epochs=3, patience=5)
    assert len(hist["loss"]) == 3


def test_glorot_limits_and_stage():
    ws = glorot_weights([10, 20], 6, 8, seed=3)
    for (W1, b1, W2, b2), p in zip(ws, [10, 20]):
        assert W1.shape == (p, 6) and W2.shape == (6, 8)
        assert np.abs(W1).max() <= np.sqrt(6 / (p + 6)) and np.abs(W2).max() <= np.sqrt(6 / (6 + 8))
        assert not b1.any() and not b2.any()
    again = glorot_weights([20], 6, 8, seed=3, subnet_ids=[1])
    np.testing.assert_array_equal(again[0][0], ws[1][0])      # keyed by the global sub-network number
    norm = np.arange(20, dtype=np.float32).reshape(4, 5)
    X, Y = stage(norm, [np.array([4, 0])], np.array([[1, 2]]), np.array([3, 1]))
    np.testing.assert_array_equal(X[0], [[19, 15], [9, 5]])
    np.testing.assert_array_equal(Y[0], [[16, 17], [6, 7]])


def test_fp32_oracle_tracks_fp64_oracle():
    """Calibrates the parity tolerance: fp32 vs fp64 oracle after a few epochs on the same data."""
    rng = np.random.default_rng(5)
    n, P, H, O, B = 200, 30, 16, 24, 32
    X = [rng.gamma(1.0, 1.0, size=(n, P)).astype(np.float32)]
    Y = [np.log1p(rng.poisson(2.0, size=(n, O))).astype(np.float32)]
    a = OracleNet([P], H, O, learning_rate=1e-3, batch_size=B, seed=9)
    b = OracleNet([P], H, O, learning_rate=1e-3, batch_size=B, seed=9, dtype=torch.float64)
    step_a = step_b = 0
    for e in range(3):
        perm = epoch_permutation(9, e, n)
        _, step_a = a.train_epoch(X, Y, perm, step_a)
        _, step_b = b.train_epoch(X, Y, perm, step_b)
    ya, yb = a.forward(X)[0], b.forward(X)[0]
    assert np.max(np.abs(ya - yb) / (np.abs(yb) + 1e-3)) < 1e-3


# ---- whole trajectory against two third-party pieces: torch autograd + scikit-learn's Adam ------------------------
def test_trajectory_matches_autograd_plus_sklearn_adam():
    """Ten optimiser steps of one sub-network rebuilt from parts the oracle does not contain: the loss written with
    torch.nn.functional and differentiated by autograd, the update done by scikit-learn's ``AdamOptimizer`` -- an
    independent implementation of the rule TensorFlow's Adam applies (``lr_t = lr sqrt(1-b2^t)/(1-b1^t)``,
    ``w -= lr_t m / (sqrt(v) + eps)``: epsilon outside the bias correction), which is what Keras' ``Adam`` of
    multinet.py:164 runs.  The oracle's hand-written backward pass and update must land on the same weights."""
    from sklearn.neural_network._stochastic_optimizers import AdamOptimizer
    import torch.nn.functional as F

    rng = np.random.default_rng(5)
    P, H, O, B, lr = 7, 6, 4, 16, 3e-3
    net = OracleNet([P], H, O, learning_rate=lr, batch_size=B, dropout_rate=0.0, seed=9, dtype=torch.float64)
    params = [torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in net.get_weights()[0]]
    b1, b2, eps = float(np.float32(0.9)), float(np.float32(0.999)), float(np.float32(1e-7))     # Keras stores them as fp32
    arrays = [p.detach().numpy() for p in params]                   # updated in place by scikit-learn, shared with torch
    opt = AdamOptimizer(arrays, learning_rate_init=lr, beta_1=b1, beta_2=b2, epsilon=eps)
    for step in range(10):
        x = rng.gamma(1.0, 1.0, size=(B, P))
        y = np.log1p(rng.poisson(2.0, size=(B, O))).astype(np.float64)
        xt, yt = torch.tensor(x), torch.tensor(y)
        W1, c1, W2, c2 = params
        yhat = F.softplus(F.relu(xt @ W1 + c1) @ W2 + c2)
        loss = (yt * (yt - yhat) ** 2).mean()                       # wMSE, multinet.py:36-41
        grads = torch.autograd.grad(loss, params)
        got = net.train_step([x], [y], step)
        assert got == pytest.approx(float(loss), rel=1e-12)
        opt.update_params(arrays, [g.numpy() for g in grads])
        for k in range(4):
            np.testing.assert_allclose(net.w[0][k].numpy(), arrays[k], rtol=1e-9, atol=1e-13)
