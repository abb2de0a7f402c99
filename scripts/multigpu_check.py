"""Run under torchrun on N GPUs: the sharded MultiNet (sub-networks split over ranks, NCCL all-reduce of the epoch losses,
NCCL all-gather of the prediction blocks, fused imputation tail on every rank) against the unsharded one on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py
"""
import contextlib
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pandas as pd

from deepimpute_b200 import parallel
from deepimpute_b200.multinet import MultiNet

ctx = parallel.init()
z = np.load(os.path.join(ROOT, "tests", "golden", "test_counts.npz"))
raw = pd.DataFrame(z["counts"].astype(np.float64), index=z["cells"].astype(object), columns=z["genes"].astype(object))
kw = dict(seed=1234, ncores=1, max_epochs=4, patience=100, verbose=0)
quiet = io.StringIO()
with contextlib.redirect_stdout(quiet):
    net = MultiNet(shard=ctx, **kw)
    net.fit(raw)
    out = net.predict(raw)
    host = MultiNet(shard=ctx, postprocess="host", **kw)
    host.fit(raw)
    out_host = host.predict(raw)
print("rank {}: owns sub-networks {} of {}, loss {}".format(ctx.rank, net._owned[ctx.rank], len(net.predictors),
                                                            ["%.6f" % v for v in net.history["loss"]]), flush=True)
np.testing.assert_allclose(out.values, out_host.values, rtol=1e-12)      # fused tail == numpy tail on gathered blocks
ctx.barrier()
if ctx.rank == 0:
    with contextlib.redirect_stdout(quiet):
        one = MultiNet(device=0, **kw)
        one.fit(raw)
        ref = one.predict(raw)
    np.testing.assert_allclose(net.history["loss"], one.history["loss"], rtol=1e-6)
    np.testing.assert_allclose(net.history["val_loss"], one.history["val_loss"], rtol=1e-6)
    np.testing.assert_array_equal(out.values, ref.values)                # per-sub-network results do not depend on the sharding
    assert abs(net.test_metrics["MSE"] - one.test_metrics["MSE"]) <= 1e-9 * one.test_metrics["MSE"]
    print("multigpu_check OK: world {} == single GPU (imputed matrix bit-identical, losses within 1e-6)".format(ctx.world_size))
ctx.barrier()
import torch.distributed as dist
if dist.is_initialized():
    dist.destroy_process_group()
