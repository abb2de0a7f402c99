"""Scratch: localise a failing kernel (run with DEEPIMPUTE_B200_DEBUG_SYNC=1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from deepimpute_b200.engine import Engine, epoch_permutation
from test_engine_gpu import make_problem

mode = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
order = sys.argv[2] if len(sys.argv) > 2 else "pve"
size = sys.argv[3] if len(sys.argv) > 3 else "small"
if len(sys.argv) > 4 and sys.argv[4] == "torch":
    import torch
    torch.zeros(8, device="cuda").sum().item()
if size == "small":
    n_pred, H, O, B, N, G = [90, 41], 40, 64, 32, 240, 900
else:
    n_pred, H, O, B, N, G = [500, 523, 480, 512], 256, 512, 64, 2000, 5000
norm, pred_idx, targ_idx = make_problem(N, G, n_pred, O, seed=5)
eng = Engine(n_pred, hidden=H, sub_outputdim=O, learning_rate=5e-4, batch_size=B, seed=7, math_mode=mode)
eng.set_data(norm, pred_idx, targ_idx)
cells = np.random.default_rng(1).permutation(N)
nt = N // 20
eng.set_split(np.sort(cells[nt:]).astype(np.int32), cells[:nt].astype(np.int32))
for ch in order:
    if ch == "p":
        print("predict", eng.predict().sum(), flush=True)
    if ch == "v":
        print("val", eng.validation_loss(), flush=True)
    if ch == "e":
        print("epoch", eng.train_epoch(epoch_permutation(7, 0, N - nt)), flush=True)
    if ch == "s":
        print("step", eng.train_step(np.arange(B, dtype=np.int32)), flush=True)
