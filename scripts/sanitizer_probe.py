"""Small end-to-end run for compute-sanitizer: every kernel family of the tf32x3 engine (converter-warp / TMA-only with
and without split K), the fp32 kernels, epoch graph + dependent launches, validation, predict, the fused impute tail,
gene statistics and predictor selection -- on a problem small enough to finish under memcheck / racecheck."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from deepimpute_b200 import partition
from deepimpute_b200.engine import Engine, epoch_permutation

rng = np.random.default_rng(0)
N, G, O, H, B = 300, 700, 64, 48, 64
n_pred = [70, 33, 96]
raw = rng.poisson(rng.gamma(0.6, 3.0, size=(1, G)), size=(N, G)).astype(np.float32)
perm = rng.permutation(G)
targ = perm[:3 * O].reshape(3, O).astype(np.int32)
pred = [rng.choice(perm[3 * O:], p, replace=False).astype(np.int32) for p in n_pred]
tr, te = np.arange(0, 250, dtype=np.int32), np.arange(250, N, dtype=np.int32)
cases = [("tf32x3", {"DEEPIMPUTE_B200_LT": "0"}), ("tf32x3", {"DEEPIMPUTE_B200_LT": "1", "DEEPIMPUTE_B200_SPLITK": "1"}),
         ("tf32x3", {"DEEPIMPUTE_B200_LT": "1", "DEEPIMPUTE_B200_SPLITK": "2"}), ("tf32", {}), ("fp32", {})]
for mode, env in cases:
    for k in ("DEEPIMPUTE_B200_LT", "DEEPIMPUTE_B200_SPLITK"):
        os.environ.pop(k, None)
    os.environ.update(env)
    eng = Engine(n_pred, hidden=H, sub_outputdim=O, batch_size=B, seed=3, math_mode=mode)
    eng.set_counts(raw, pred, targ)
    eng.set_split(tr, te)
    for ep in range(2):
        loss, val = eng.train_epoch(epoch_permutation(3, ep, len(tr)))
    step_loss = eng.train_step(np.arange(40, dtype=np.int32))
    out = eng.predict()
    imp = eng.impute(policy="restore", dtype=np.float32)
    print(mode, env, eng.describe(), "loss %.5f val %.5f step %.5f pred %.4f imputed %.4f fallbacks %d"
          % (loss, val, step_loss, float(out.mean()), float(imp.mean()), eng.graph_fallbacks()), flush=True)
    eng.close()
mean, var, _ = partition.gene_stats_gpu(raw)
picked, _ = partition.choose_predictors_gpu(raw, targ, np.arange(G), np.array(["g%04d" % j for j in range(G)], dtype=object), 5)
print("gene stats", float(mean.sum()), float(var.sum()), "predictors", [len(p) for p in picked])
