"""Which knob of the converter-warp kernel family costs accuracy?  Trains the 3-sub-network problem of
tests/test_benchmarked_parity_gpu.py::test_kernel_families_follow_the_oracle for 2 epochs under several knob settings and
prints the error of predictions and weights against the CPU oracle (max |a - b| / max |b|, and the 99.9th percentile)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from deepimpute_b200.engine import Engine, epoch_permutation
from oracle.multinet_oracle import OracleNet, stage

H, O, RATE, SEED = 256, 512, 0.2, 1234
rng = np.random.default_rng(11)
n_pred = [540, 513, 600]
BATCH = int(os.environ.get("FA_BATCH", "64"))
N, G = 40 * BATCH - 17 + 128, 2400
lam = rng.gamma(0.6, 3.0, size=(1, G)) * rng.gamma(2.0, 0.5, size=(N, 1))
norm = np.log1p(rng.poisson(lam)).astype(np.float32)
perm = rng.permutation(G)
targ = perm[:3 * O].reshape(3, O).astype(np.int32)
pred_idx = [rng.choice(perm[3 * O:], p, replace=False).astype(np.int32) for p in n_pred]
tr, te = np.arange(N - 128, dtype=np.int32), np.arange(N - 128, N, dtype=np.int32)
EPOCHS = int(os.environ.get("FA_EPOCHS", "2"))
LR = float(os.environ.get("FA_LR", "1e-3"))

ref = OracleNet(n_pred, H, O, learning_rate=LR, batch_size=BATCH, dropout_rate=RATE, seed=SEED)
Xtr, Ytr = stage(norm, pred_idx, targ, tr)
step = 0
for epoch in range(EPOCHS):
    _, step = ref.train_epoch(Xtr, Ytr, epoch_permutation(SEED, epoch, len(tr)), step)
want = np.hstack(ref.forward(stage(norm, pred_idx, targ, np.arange(N))[0]))
want_w = ref.get_weights()

KNOBS = ["DEEPIMPUTE_B200_LT", "DEEPIMPUTE_B200_SPLITK", "DEEPIMPUTE_B200_TS", "DEEPIMPUTE_B200_PDL", "DEEPIMPUTE_B200_GRAPH",
         "DEEPIMPUTE_B200_PDL_PREFETCH", "DEEPIMPUTE_B200_ADAM", "DEEPIMPUTE_B200_EXPERIMENT", "DEEPIMPUTE_B200_ADAM_VEC", "DEEPIMPUTE_B200_GROUPS", "DEEPIMPUTE_B200_MATH"]
CASES = [
    ("lt", dict(DEEPIMPUTE_B200_LT="1", DEEPIMPUTE_B200_SPLITK="1")),
    ("lt split4", dict(DEEPIMPUTE_B200_LT="1", DEEPIMPUTE_B200_SPLITK="4")),
    ("conv ts", dict(DEEPIMPUTE_B200_LT="0")),
    ("conv + exact fp32 dW", dict(DEEPIMPUTE_B200_LT="0", DEEPIMPUTE_B200_EXPERIMENT="2")),
    ("conv smem", dict(DEEPIMPUTE_B200_LT="0", DEEPIMPUTE_B200_TS="0")),
    ("conv ts nopdl", dict(DEEPIMPUTE_B200_LT="0", DEEPIMPUTE_B200_PDL="0")),
    ("conv ts noprefetch", dict(DEEPIMPUTE_B200_LT="0", DEEPIMPUTE_B200_PDL_PREFETCH="0")),
    ("conv ts nograph", dict(DEEPIMPUTE_B200_LT="0", DEEPIMPUTE_B200_GRAPH="0")),
    ("conv ts 1group", dict(DEEPIMPUTE_B200_LT="0", DEEPIMPUTE_B200_GROUPS="1")),
    ("conv ts ring", dict(DEEPIMPUTE_B200_LT="0", DEEPIMPUTE_B200_ADAM="ring")),
    ("fp32", dict(DEEPIMPUTE_B200_MATH="fp32")),
]
for name, env in CASES:
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(env)
    eng = Engine(n_pred, hidden=H, sub_outputdim=O, learning_rate=LR, batch_size=BATCH, dropout_rate=RATE, seed=SEED)
    eng.set_data(norm, pred_idx, targ)
    eng.set_split(tr, te)
    t0 = time.perf_counter()
    for epoch in range(EPOCHS):
        eng.train_epoch(epoch_permutation(SEED, epoch, len(tr)))
    dt = time.perf_counter() - t0
    got = eng.predict()
    err = np.abs(got.astype(np.float64) - want) / np.abs(want).max()
    werr = 0.0
    for gw, rw in zip(eng.get_weights(), want_w):
        for a, b in zip(gw, rw):
            werr = max(werr, float(np.max(np.abs(a.astype(np.float64) - b)) / np.max(np.abs(b))))
    print("{:22s} pred max {:.2e}  p99.9 {:.2e}  median {:.2e} | weights max {:.2e} | {:.0f} ms | {}".format(
        name, err.max(), np.quantile(err, 0.999), np.median(err), werr, dt * 1e3, eng.describe()), flush=True)
    eng.close()
