"""Where do the tensor-core kernels leave the fp32 trajectory?  Same problem as scripts/family_accuracy.py, fp32 CUDA-core
engine against the default engine, step by step: worst element of every weight / moment tensor, with its moments."""
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
LR = float(os.environ.get("LR", "1e-3"))
order = epoch_permutation(SEED, 0, N - 128)

def make(mode):
    e = Engine(n_pred, hidden=H, sub_outputdim=O, learning_rate=LR, batch_size=64, dropout_rate=RATE, seed=SEED, math_mode=mode)
    e.set_data(norm, pred_idx, targ)
    return e

a, b = make("fp32"), make(os.environ.get("MODE", "tf32x3"))
names = ["W1", "b1", "W2", "b2"]
checkpoints = [1, 2, 5, 10, 20, 39]
for step in range(39):
    rows = order[step * 64:(step + 1) * 64]
    la, lb = a.train_step(rows, step), b.train_step(rows, step)
    if step + 1 in checkpoints:
        print("step {:3d}: loss fp32 {:.7f} tc {:.7f} (rel {:.1e})".format(step + 1, la, lb, abs(la - lb) / la))
        for s in range(3):
            wa, wb = a.get_weights()[s], b.get_weights()[s]
            (ma, _), (mb, _) = a.get_adam_state(s), b.get_adam_state(s)
            for k, nm in enumerate(names):
                d = np.abs(wa[k].astype(np.float64) - wb[k])
                i = np.unravel_index(np.argmax(d), d.shape)
                m_a, v_a, m_b, v_b = ma[2 * k][i], ma[2 * k + 1][i], mb[2 * k][i], mb[2 * k + 1][i]
                dm = np.abs(ma[2 * k].astype(np.float64) - mb[2 * k])
                print("   net {} {:2s}: max|dw| {:.2e} ({:.1f} lr) at {} w {:+.5e}/{:+.5e} m {:+.3e}/{:+.3e} sqrt(v) {:.3e}/{:.3e} | max|dm| {:.2e} of scale {:.2e}"
                      .format(s, nm, d.max(), d.max() / LR, i, wa[k][i], wb[k][i], m_a, m_b, np.sqrt(v_a), np.sqrt(v_b), dm.max(),
                              np.abs(ma[2 * k]).max()))
