"""Stand-in for ``deepimpute_b200.engine.Engine`` built on the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Lets the host-side logic of ``MultiNet`` (partitioning, split, sharding, post-processing) be exercised on a box
without a GPU, including world-size-2 gloo runs.  The product never imports this."""
import numpy as np

from oracle.multinet_oracle import OracleNet, epoch_permutation, stage
from oracle.postprocess_oracle import impute_tail, log1p_norm


class _History:
    def __init__(self):
        self.history = {"loss": [], "val_loss": []}


class FakeEngine:
    def __init__(self, inputdims, hidden=256, sub_outputdim=512, learning_rate=1e-4, batch_size=64,
                 dropout_rate=0.2, seed=1234, math_mode=None, device=None, subnet_ids=None, **_):
        self.n_pred = list(inputdims)
        self.S, self.H, self.O, self.B, self.seed = len(self.n_pred), hidden, sub_outputdim, batch_size, seed
        self.subnet_ids = list(range(self.S)) if subnet_ids is None else list(subnet_ids)
        self.net = OracleNet(self.n_pred, hidden, sub_outputdim, learning_rate, batch_size, dropout_rate, seed,
                             subnet_ids=self.subnet_ids)
        self.saved = None

    def set_data(self, norm, pred_idx, targ_idx):
        self.norm, self.pred_idx, self.targ_idx = np.asarray(norm, np.float32), pred_idx, np.asarray(targ_idx)
        self.n_cells = self.norm.shape[0]

    def set_counts(self, raw, pred_idx, targ_idx):
        self.raw = np.asarray(raw)
        self.set_data(log1p_norm(self.raw), pred_idx, targ_idx)

    def impute(self, policy="restore", pred=None, slot_gene=None, **_):
        if pred is None:
            pred, slot_gene = self.predict(), np.asarray(self.targ_idx).reshape(-1)
        return impute_tail(self.raw, np.asarray(pred), slot_gene, policy)

    def fit(self, train_rows, test_rows, epochs, patience=5, verbose=0, perm_fn=None, on_epoch_end=None):
        Xtr, Ytr = stage(self.norm, self.pred_idx, self.targ_idx, train_rows)
        Xte, Yte = stage(self.norm, self.pred_idx, self.targ_idx, test_rows)
        hist, best, wait, step = _History(), np.inf, 0, 0
        for e in range(epochs):
            loss, step = self.net.train_epoch(Xtr, Ytr, epoch_permutation(self.seed, e, len(train_rows)), step)
            val = self.net.loss(Xte, Yte)
            if on_epoch_end is not None:
                loss, val = on_epoch_end(e, loss, val)
            hist.history["loss"].append(loss)
            hist.history["val_loss"].append(val)
            if val < best:
                best, wait = val, 0
            else:
                wait += 1
                if wait >= patience:
                    break
        return hist

    def predict(self, rows=None):
        rows = np.arange(self.n_cells) if rows is None else rows
        X, _ = stage(self.norm, self.pred_idx, self.targ_idx, rows)
        return np.hstack(self.net.forward(X)).astype(np.float32)

    predict_block = predict

    def get_weights(self):
        return self.net.get_weights()

    def save(self, path, **extra):
        self.saved = (path, extra)
