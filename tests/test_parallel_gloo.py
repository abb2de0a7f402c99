"""N > 1 path on CPU: two gloo processes each own half of the sub-networks (oracle-backed stand-in engine), exchange
the per-epoch losses and all-gather the predicted blocks; the result must equal the single-process run."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, synthetic_counts
from deepimpute_b200.parallel import ShardContext, assign_subnets

WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from conftest import synthetic_counts
from deepimpute_b200 import MultiNet, parallel
from fake_engine import FakeEngine

class CpuMultiNet(MultiNet):
    def _make_engine(self, inputdims, **kw):
        return FakeEngine(inputdims, **kw)

ctx = parallel.init(backend="gloo")
raw = synthetic_counts(100, 80, seed=2)
net = CpuMultiNet(ncores=1, sub_outputdim=16, max_epochs=4, patience=2, seed=3, verbose=0, shard=ctx,
                  architecture=[dict(type="dense", neurons=10, activation="relu"), dict(type="dropout", rate=0.2)])
net.fit(raw, NN_lim=40, minVMR=0.0)
out = net.predict(raw)
np.savez({out!r} + ".rank{{}}.npz".format(ctx.rank), imputed=out.values, loss=net.history["loss"],
         val=net.history["val_loss"], corr=net.test_metrics["correlation"], owned=np.asarray(net._owned, dtype=object)
         if False else np.asarray([len(o) for o in net._owned or [[0]]]))
ctx.barrier()
"""


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_world(tmp_path, world):
    out = str(tmp_path / "w{}".format(world))
    script = tmp_path / "worker{}.py".format(world)
    script.write_text(WORKER.format(root=ROOT, out=out))
    port = str(free_port())
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=port, OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate(timeout=600)[0] for p in procs]
    for p, log in zip(procs, logs):
        assert p.returncode == 0, log[-3000:]
    return [np.load("{}.rank{}.npz".format(out, r)) for r in range(world)]


def test_two_gloo_ranks_match_one(tmp_path):
    one = run_world(tmp_path, 1)[0]
    two = run_world(tmp_path, 2)
    assert list(two[0]["owned"]) == [2, 2]
    for r in two:
        # same early-stopping trajectory on every rank, equal to the unsharded run (summed losses)
        np.testing.assert_allclose(r["loss"], one["loss"], rtol=1e-6)
        np.testing.assert_allclose(r["val"], one["val"], rtol=1e-6)
        np.testing.assert_allclose(r["imputed"], one["imputed"], rtol=1e-6)
        assert float(r["corr"]) == pytest.approx(float(one["corr"]), rel=1e-6)


def test_assign_subnets_balances_parameter_counts():
    n_pred = [2000, 600, 610, 590, 1500, 620, 605, 598, 640, 2560]
    for world in (1, 2, 4, 8):
        owned = assign_subnets(n_pred, world)
        assert sorted(s for o in owned for s in o) == list(range(len(n_pred)))
        cost = [sum(n_pred[s] * 256 + 131072 for s in o) for o in owned]
        assert max(cost) - min(cost) <= max(n_pred) * 256 + 131072
    assert assign_subnets([5, 5], 4)[2:] == [[], []]


def test_single_rank_context_is_a_no_op():
    ctx = ShardContext()
    assert not ctx.distributed and ctx.sum_scalars(1.5, 2.5) == (1.5, 2.5)
    block = np.arange(12, dtype=np.float32).reshape(3, 4)
    np.testing.assert_array_equal(ctx.gather_blocks(block, [[0, 1]], 2), block)

