"""Scratch: c3-shaped explicit steps for the in-kernel pipeline trace, or a c2-sized upload_counts + impute pass.

    DEEPIMPUTE_B200_TRACE=1 python scripts/trace_step.py step tf32x3     # trace of CTA (0,0,0) of each kernel on stderr
    python scripts/trace_step.py impute                                  # 10k x 5k counts: log1p + fused tail (for ncu)
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from deepimpute_b200.engine import Engine

what = sys.argv[1] if len(sys.argv) > 1 else "step"
mode = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
rng = np.random.default_rng(0)
if what == "step":
    S, H, O, B, N, G = int(os.environ.get("TRACE_S", "40")), 256, 512, 64, 512, 24000
    n_pred = [int(p) for p in rng.integers(512, 561, size=S)]
    norm = np.log1p(rng.poisson(2.0, size=(N, G))).astype(np.float32)
    perm = rng.permutation(G)
    targ = perm[:S * O].reshape(S, O).astype(np.int32)
    pred_idx = [rng.choice(G, p, replace=False).astype(np.int32) for p in n_pred]
    eng = Engine(n_pred, hidden=H, sub_outputdim=O, batch_size=B, seed=1, math_mode=mode)
    eng.set_data(norm, pred_idx, targ)
    for i in range(3):
        print("step", i, eng.train_step(np.arange(i * B, (i + 1) * B, dtype=np.int32)), file=sys.stderr, flush=True)
else:
    N, G, S, O = 10000, 5000, 9, 512
    raw = rng.poisson(rng.gamma(0.5, 4.0, size=(1, G)), size=(N, G)).astype(np.float32)
    perm = rng.permutation(G)
    targ = perm[:S * O].reshape(S, O).astype(np.int32)
    targ[-1, -40:] = targ[0, :40]
    pred_idx = [rng.choice(G, 540, replace=False).astype(np.int32) for _ in range(S)]
    eng = Engine([540] * S, batch_size=64, seed=1, math_mode=mode)
    for rep in range(2):
        t0 = time.perf_counter()
        eng.set_counts(raw, pred_idx, targ)
        t1 = time.perf_counter()
        out = eng.impute()
        t2 = time.perf_counter()
        print("upload_counts %.1f ms, impute %.1f ms, checksum %.6e" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, out.sum()))
