"""Inference-only driver for ncu: S sub-networks of the default topology over N random cells (c3-like shapes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from deepimpute_b200.engine import Engine

S, P, N, G = int(os.environ.get("S", 40)), 544, int(os.environ.get("N", 16384)), 8192
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
rng = np.random.default_rng(0)
norm = np.log1p(rng.poisson(2.0, size=(N, G))).astype(np.float32)
pred_idx = [rng.choice(G, P, replace=False).astype(np.int32) for _ in range(S)]
targ_idx = rng.integers(0, G, size=(S, 512)).astype(np.int32)
eng = Engine([P] * S, math_mode=mode)
eng.set_data(norm, pred_idx, targ_idx)
import torch
out = torch.empty((N, S * 512), dtype=torch.float32, device="cuda")
for _ in range(3):
    eng.predict_device(out.data_ptr(), S * 512)
    print("predict device ms", eng.last_device_ms(), "TFLOP/s (useful fp32-equivalent)",
          2.0 * N * S * (P * 256 + 256 * 512) / (eng.last_device_ms() * 1e-3) / 1e12, flush=True)
