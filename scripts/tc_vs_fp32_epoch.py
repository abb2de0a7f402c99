"""Epoch path vs step path, tensor-core engine vs fp32 engine (see scripts/tc_vs_fp32.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from deepimpute_b200.engine import Engine, epoch_permutation

H, O, RATE, SEED = 256, 512, 0.2, 1234
rng = np.random.default_rng(11)
n_pred = [540, 513, 600]
N, G = 40 * 64 - 17 + 128, 2400
lam = rng.gamma(0.6, 3.0, size=(1, G)) * rng.gamma(2.0, 0.5, size=(N, 1))
norm = np.log1p(rng.poisson(lam)).astype(np.float32)
perm = rng.permutation(G)
targ = perm[:3 * O].reshape(3, O).astype(np.int32)
pred_idx = [rng.choice(perm[3 * O:], p, replace=False).astype(np.int32) for p in n_pred]
LR = 1e-3
names = ["W1", "b1", "W2", "b2"]


def make(mode):
    e = Engine(n_pred, hidden=H, sub_outputdim=O, learning_rate=LR, batch_size=64, dropout_rate=RATE, seed=SEED, math_mode=mode)
    e.set_data(norm, pred_idx, targ)
    return e


def report(tag, a, b):
    print(tag)
    for s in range(3):
        wa, wb = a.get_weights()[s], b.get_weights()[s]
        (ma, _), (mb, _) = a.get_adam_state(s), b.get_adam_state(s)
        for k, nm in enumerate(names):
            d = np.abs(wa[k].astype(np.float64) - wb[k])
            i = np.unravel_index(np.argmax(d), d.shape)
            big = int((d > 10 * LR * 1e-2).sum())
            print("   net {} {:2s}: max|dw| {:.2e} ({:.2f} lr) at {} of shape {} | {} elements differ by > 0.1 lr | m {:+.3e}/{:+.3e} sqrt(v) {:.3e}/{:.3e}"
                  .format(s, nm, d.max(), d.max() / LR, tuple(int(x) for x in i), wa[k].shape, big,
                          ma[2 * k][i], mb[2 * k][i], np.sqrt(ma[2 * k + 1][i]), np.sqrt(mb[2 * k + 1][i])))
            if nm == "W1" and big:
                rows = np.unique(np.nonzero(d > 10 * LR * 1e-2)[0])
                print("        rows (predictors) involved:", rows[:20], "... of", len(rows), "| columns:", np.unique(np.nonzero(d > 10 * LR * 1e-2)[1])[:20])
                x = norm[:, pred_idx[s][rows[:5]]]
                print("        those predictors: fraction of non-zero cells", (x > 0).mean(0))


n_tr = N - 128
tr, te = np.arange(n_tr, dtype=np.int32), np.arange(n_tr, N, dtype=np.int32)

def pred_err(a, b):
    pa, pb = a.predict(), b.predict()
    err = np.abs(pa.astype(np.float64) - pb) / np.abs(pa).max()
    print("    predictions: max {:.2e} p99.9 {:.2e} median {:.2e}".format(err.max(), np.quantile(err, 0.999), np.median(err)))

# (E) two epochs through di_train_epoch
a, b = make("fp32"), make("tf32x3")
for e in (a, b):
    e.set_split(tr, te)
for ep in range(2):
    order = epoch_permutation(SEED, ep, n_tr)
    la, lb = a.train_epoch(order), b.train_epoch(order)
    print("epoch", ep, "loss/val fp32", la, "tc", lb)
    report("(E) after epoch {} (di_train_epoch)".format(ep + 1), a, b)
    pred_err(a, b)
a.close(); b.close()

# (F) epoch 1 through di_train_epoch, epoch 2 as explicit steps
a, b = make("fp32"), make("tf32x3")
for e in (a, b):
    e.set_split(tr, te)
    e.train_epoch(epoch_permutation(SEED, 0, n_tr))
order = epoch_permutation(SEED, 1, n_tr)
for step in range(40):
    rows = order[step * 64:(step + 1) * 64]
    a.train_step(rows, 40 + step); b.train_step(rows, 40 + step)
report("(F) epoch 2 as explicit steps", a, b)
pred_err(a, b)
a.close(); b.close()

# (G) two epochs, fp32 engine against itself: di_train_epoch vs explicit steps (is the EPOCH path the odd one?)
a, b = make("fp32"), make("fp32")
a.set_split(tr, te)
for ep in range(2):
    order = epoch_permutation(SEED, ep, n_tr)
    a.train_epoch(order)
    for step in range(40):
        b.train_step(order[step * 64:(step + 1) * 64], 40 * ep + step)
report("(G) fp32: epochs vs explicit steps", a, b)
a.close(); b.close()
a, b = make("tf32x3"), make("tf32x3")
a.set_split(tr, te)
for ep in range(2):
    order = epoch_permutation(SEED, ep, n_tr)
    a.train_epoch(order)
    for step in range(40):
        b.train_step(order[step * 64:(step + 1) * 64], 40 * ep + step)
report("(H) tf32x3: epochs vs explicit steps", a, b)
a.close(); b.close()
