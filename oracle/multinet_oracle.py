"""CPU restatement of the neural-network hot path of DeepImpute.  TEST INFRASTRUCTURE (see oracle/__init__.py).

What is restated, and from where:

* topology ``Input(P_s) -> Dense(H, relu) -> Dropout(r) -> Dense(O, softplus)`` per sub-network, no shared weights:
  reference ``deepimpute/multinet.py:99-103`` (default architecture) and ``:132-148`` (graph);
* loss ``wMSE = mean(y * (y - yhat)^2)`` over batch x O, summed over sub-networks: ``multinet.py:36-41``; the sum is
  Keras' rule for multi-output models [upstream-Keras 2.x ``Model.compile``: total loss = sum of output losses];
* optimiser ``Adam(lr)``: ``multinet.py:164``; update rule of TensorFlow's ``ResourceApplyAdam`` kernel
  [upstream-TF 2.x ``training_ops.cc``]: ``lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g*g-v)(1-b2);
  w -= lr_t*m/(sqrt(v)+eps)`` with b1 .9, b2 .999, eps 1e-7, one step counter t shared by all variables;
* training loop ``model.fit(..., validation_data, epochs, batch_size, callbacks=[EarlyStopping('val_loss',
  patience)])``: ``multinet.py:238-244``; [upstream-Keras] shuffle every epoch, partial last batch kept, logged
  ``loss`` = sample-weighted mean of batch losses, validation in inference mode, EarlyStopping(min_delta=0,
  restore_best_weights=False) stops after ``patience`` epochs without a new best;
* inference ``model.predict`` + ``np.hstack``: ``multinet.py:253``, ``:278-280``;
* data staging ``X_s = norm[cells, predictors_s]``, ``Y_s = norm[cells, targets_s]``: ``multinet.py:231-235``.

Randomness that TensorFlow would draw (initial weights, epoch shuffles, dropout masks) cannot be matched
without TensorFlow; it is *defined* instead so that the CUDA path can reproduce it exactly:
Glorot-uniform kernels / zero biases [upstream-Keras Dense defaults] from ``numpy.random.default_rng([seed, s])``,
epoch permutations from ``default_rng([seed, 0x5eed, epoch])``, dropout masks from Philox (``oracle/philox.py``).
"""
import numpy as np
import torch

from .philox import dropout_keep_mask


def glorot_weights(n_pred, hidden, out, seed, subnet_ids=None):
    """[upstream-Keras] Dense defaults: kernel ~ U(+-sqrt(6/(fan_in+fan_out))), bias = 0.  One rng per sub-network,
    keyed by its global number."""
    ws = []
    for s, p in zip(range(len(n_pred)) if subnet_ids is None else subnet_ids, n_pred):
        rng = np.random.default_rng([int(seed), s])
        l1 = np.sqrt(6.0 / (p + hidden))
        l2 = np.sqrt(6.0 / (hidden + out))
        ws.append((rng.uniform(-l1, l1, size=(p, hidden)).astype(np.float32), np.zeros(hidden, np.float32),
                   rng.uniform(-l2, l2, size=(hidden, out)).astype(np.float32), np.zeros(out, np.float32)))
    return ws


def epoch_permutation(seed, epoch, n):
    return np.random.default_rng([int(seed), 0x5EED, int(epoch)]).permutation(n).astype(np.int32)


def _round_operand(x, mode):
    """Emulate tensor-core operand precision: 'tf32' keeps 10 mantissa bits (truncation), 'bf16' 7 (truncation)."""
    if mode is None:
        return x
    bits = {"tf32": 13, "bf16": 16}[mode]
    i = x.to(torch.float32).contiguous().view(torch.int32)
    i = i & ~((1 << bits) - 1)
    return i.view(torch.float32).to(x.dtype)


class OracleNet:
    """S independent two-layer perceptrons trained together, Keras semantics (see module docstring)."""

    def __init__(self, n_pred, hidden, out, learning_rate=1e-4, batch_size=64, dropout_rate=0.2, seed=1234,
                 beta1=0.9, beta2=0.999, epsilon=1e-7, dtype=torch.float32, operand_round=None, subnet_ids=None,
                 mask_mode="philox"):
        self.n_pred, self.H, self.O = list(n_pred), hidden, out
        self.S = len(self.n_pred)
        self.lr, self.B, self.rate, self.seed = learning_rate, batch_size, dropout_rate, seed
        # di_config carries these as float32; use the same values
        self.b1, self.b2, self.eps = (float(np.float32(x)) for x in (beta1, beta2, epsilon))
        self.dtype = dtype
        self.round = operand_round
        # "philox": the defined, GPU-reproducible mask (parity tests).  "torch": a plain torch.rand mask, as cheap as
        # TensorFlow's own dropout RNG -- used only when the oracle is TIMED as the CPU baseline (bench.py), where
        # numpy-Philox would charge the baseline for an RNG cost the reference does not pay.
        self.mask_mode = mask_mode
        self.t = 0
        self.ids = list(range(self.S)) if subnet_ids is None else [int(s) for s in subnet_ids]
        self.set_weights(glorot_weights(self.n_pred, hidden, out, seed, self.ids))

    # -- state ------------------------------------------------------------------------------------------
    def set_weights(self, weights):
        self.w = [[torch.as_tensor(np.array(a, copy=True)).to(self.dtype) for a in ws] for ws in weights]
        self.m = [[torch.zeros_like(a) for a in ws] for ws in self.w]
        self.v = [[torch.zeros_like(a) for a in ws] for ws in self.w]
        self.t = 0

    def get_weights(self):
        return [tuple(a.numpy().copy() for a in ws) for ws in self.w]

    def _keep_scale(self):
        if self.dtype == torch.float32:
            return float(np.float32(1.0) / (np.float32(1.0) - np.float32(self.rate)))
        return 1.0 / (1.0 - self.rate)

    def _mm(self, a, b):
        return _round_operand(a, self.round) @ _round_operand(b, self.round)

    # -- forward / backward -----------------------------------------------------------------------------
    def _forward_one(self, s, x, training, step):
        W1, b1, W2, b2 = self.w[s]
        z1 = self._mm(x, W1) + b1
        a = torch.relu(z1)
        if training and self.rate > 0:
            if self.mask_mode == "philox":
                keep = torch.from_numpy(dropout_keep_mask(self.seed, self.ids[s], step, x.shape[0], self.H, self.rate))
            else:
                keep = torch.rand(x.shape[0], self.H) >= self.rate
            h = torch.where(keep, a * self._keep_scale(), torch.zeros_like(a))
        else:
            h = a
        z2 = self._mm(h, W2) + b2
        yhat = torch.nn.functional.softplus(z2, beta=1.0, threshold=1e9) if self.dtype == torch.float64 \
            else torch.clamp(z2, min=0) + torch.log1p(torch.exp(-torch.abs(z2)))
        return z1, h, z2, yhat

    def forward(self, X_list):
        """Inference forward (model.predict): list of [n, O] arrays."""
        return [self._forward_one(s, torch.as_tensor(x).to(self.dtype), False, 0)[3].numpy()
                for s, x in enumerate(X_list)]

    def loss(self, X_list, Y_list):
        """Sum over sub-networks of mean(y (y - yhat)^2), inference mode (Keras val_loss)."""
        tot = 0.0
        for s, (x, y) in enumerate(zip(X_list, Y_list)):
            y = torch.as_tensor(y).to(self.dtype)
            yhat = self._forward_one(s, torch.as_tensor(x).to(self.dtype), False, 0)[3]
            tot += float((y * (y - yhat) ** 2).mean())
        return tot

    def gradients(self, s, x, y, step, training=True):
        """Loss of sub-network s on one batch and its gradients (dW1, db1, dW2, db2) plus intermediates."""
        x = torch.as_tensor(x).to(self.dtype)
        y = torch.as_tensor(y).to(self.dtype)
        W1, b1, W2, b2 = self.w[s]
        z1, h, z2, yhat = self._forward_one(s, x, training, step)
        n = x.shape[0]
        L = (y * (y - yhat) ** 2).mean()
        dz2 = 2.0 * y * (yhat - y) * torch.sigmoid(z2) / (n * self.O)
        dW2 = self._mm(h.T, dz2)
        db2 = dz2.sum(0)
        dh = self._mm(dz2, W2.T)
        scale = self._keep_scale() if (training and self.rate > 0) else 1.0
        dz1 = torch.where(h > 0, dh * scale, torch.zeros_like(dh))
        dW1 = self._mm(x.T, dz1)
        db1 = dz1.sum(0)
        return float(L), (dW1, db1, dW2, db2), dict(z1=z1, h=h, z2=z2, yhat=yhat, dz2=dz2, dz1=dz1)

    def train_step(self, X_list, Y_list, step):
        """One optimiser step on one batch for all sub-networks; returns the summed loss before the update."""
        self.t += 1
        lr_t = self.lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        if self.dtype == torch.float32:
            lr_t = float(np.float32(lr_t))
        total = 0.0
        for s in range(self.S):
            L, grads, _ = self.gradients(s, X_list[s], Y_list[s], step)
            total += L
            for k, g in enumerate(grads):
                m, v, w = self.m[s][k], self.v[s][k], self.w[s][k]
                m += (g - m) * (1.0 - self.b1)
                v += (g * g - v) * (1.0 - self.b2)
                w -= lr_t * m / (torch.sqrt(v) + self.eps)
        return total

    # -- epoch loop ---------------------------------------------------------------------------------------
    def train_epoch(self, X_train, Y_train, perm, first_step):
        """One epoch over rows ``perm`` in batches of B (partial last batch kept); returns Keras' logged loss."""
        n = len(perm)
        acc, step = 0.0, first_step
        for lo in range(0, n, self.B):
            rows = perm[lo:lo + self.B]
            L = self.train_step([x[rows] for x in X_train], [y[rows] for y in Y_train], step)
            acc += L * len(rows)
            step += 1
        return acc / n, step

    def fit(self, X_train, Y_train, X_test, Y_test, epochs, patience=5, perm_fn=None, verbose=0):
        """model.fit with EarlyStopping(monitor='val_loss', patience); returns {'loss': [...], 'val_loss': [...]}."""
        n = X_train[0].shape[0]
        perm_fn = perm_fn or (lambda e: epoch_permutation(self.seed, e, n))
        hist = {"loss": [], "val_loss": []}
        best, wait, step = np.inf, 0, 0
        for e in range(epochs):
            loss, step = self.train_epoch(X_train, Y_train, perm_fn(e), step)
            val = self.loss(X_test, Y_test)
            hist["loss"].append(loss)
            hist["val_loss"].append(val)
            if verbose:
                print("epoch {} loss {:.6f} val_loss {:.6f}".format(e + 1, loss, val))
            if val < best:
                best, wait = val, 0
            else:
                wait += 1
                if wait >= patience:
                    break
        return hist


def stage(norm, pred_idx, targ_idx, rows):
    """X_s = norm[rows][:, predictors_s], Y_s = norm[rows][:, targets_s] (multinet.py:231-235)."""
    sub = norm[rows]
    return [np.ascontiguousarray(sub[:, p]) for p in pred_idx], [np.ascontiguousarray(sub[:, t]) for t in targ_idx]
