"""Scratch: end-to-end drift of the tensor-core variants against the fp32 path (5 epochs on test.csv)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, pandas as pd
from deepimpute_b200 import MultiNet

z = np.load(os.path.join(ROOT, "tests", "golden", "test_counts.npz"))
raw = pd.DataFrame(z["counts"].astype(np.float64), index=z["cells"].astype(object), columns=z["genes"].astype(object))
epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 5
res = {}
for name, mode, exp in [("fp32", "fp32", "0"), ("tf32", "tf32", "0"), ("tf32x3", "tf32x3", "0"),
                        ("tf32x3, exact fp32 dW", "tf32x3", "2")]:
    os.environ["DEEPIMPUTE_B200_EXPERIMENT"] = exp
    net = MultiNet(seed=1234, ncores=1, max_epochs=epochs, patience=1000, verbose=0, math_mode=mode)
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        net.fit(raw)
        out = net.predict(raw).values
    res[name] = out
    if name != "fp32":
        zero = raw.values == 0
        rel = np.abs(out[zero] - res["fp32"][zero]) / (np.abs(res["fp32"][zero]) + 1e-3)
        print("{:26s} median {:.2e}  p99 {:.2e}  max {:.2e}".format(name, np.median(rel), np.quantile(rel, .99), rel.max()), flush=True)
