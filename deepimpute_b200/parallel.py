"""Sharding the sub-networks of one ``MultiNet`` over the GPUs of a box: one process per GPU.

The reference trains all S branches inside one Keras model on one CPU (``multinet.py:132-148``); the branches
share no weights and write disjoint output columns, so the path partitions over sub-networks with exactly two
exchange points:

* once per epoch the two scalars ``loss`` / ``val_loss`` are summed over ranks (``all_reduce``), because Keras sums
  the per-output losses and ``EarlyStopping`` watches that sum (``multinet.py:242-243``) -- every rank therefore
  takes the same stop decision as the single model would;
* once per ``predict`` the per-rank blocks ``[N, S_r * O]`` are all-gathered over NVLink (NCCL) into the full
  ``[N, S * O]`` matrix of ``np.hstack(model.predict(...))`` (``multinet.py:278-280``).

Launch with ``python -m torch.distributed.run --nproc-per-node N ...``; every rank runs the same script, calls
``init()`` and then uses ``MultiNet(..., shard=ctx)`` like the single-GPU object.  Per-sub-network randomness
(initial weights, dropout masks) is keyed by the *global* sub-network number, so results do not depend on the
number of ranks.  On CPU-only hosts the same code runs over ``gloo`` (used by the tests with a stand-in engine).
"""
import os

import numpy as np


def assign_subnets(n_pred, world_size, hidden=256, out=512):
    """Balanced assignment of sub-networks to ranks: longest-processing-time first on the parameter count
    ``P_s*H + H*O`` (the optimiser-state traffic that bounds a training step).  Returns a list of sorted lists."""
    cost = [int(p) * hidden + hidden * out for p in n_pred]
    order = sorted(range(len(cost)), key=lambda s: (-cost[s], s))
    load = [0] * world_size
    owned = [[] for _ in range(world_size)]
    for s in order:
        r = min(range(world_size), key=lambda k: (load[k], len(owned[k]), k))
        owned[r].append(s)
        load[r] += cost[s]
    return [sorted(o) for o in owned]


class ShardContext:
    """Rank/world of this process plus the two collectives the path needs."""

    def __init__(self, rank=0, world_size=1, device=None):
        self.rank, self.world_size = int(rank), int(world_size)
        self.device = device          # CUDA ordinal, or None on a CPU-only (gloo) run

    @property
    def distributed(self):
        return self.world_size > 1

    def _dist(self):
        import torch.distributed as dist
        return dist

    def _torch_device(self):
        import torch
        return torch.device("cuda", self.device) if self.device is not None else torch.device("cpu")

    def sum_scalars(self, *values):
        """Sum of each scalar over ranks (per-epoch loss / val_loss exchange)."""
        if not self.distributed:
            return values
        import torch
        t = torch.tensor(values, dtype=torch.float64, device=self._torch_device())
        self._dist().all_reduce(t)
        return tuple(float(x) for x in t.cpu())

    def gather_blocks(self, block, owned, width):
        """All-gather the per-rank prediction blocks and put the columns back in global sub-network order.

        ``block``: this rank's ``[N, len(owned[rank]) * width]`` float32 tensor (CUDA for nccl, CPU for gloo) or
        numpy array; ``owned``: the assignment of ``assign_subnets``; ``width``: O, columns per sub-network.
        Returns a float32 numpy array ``[N, S * width]``.  Ranks may own different numbers of sub-networks, so
        every block is padded to the widest one before the single ``all_gather``.
        """
        import torch
        if isinstance(block, np.ndarray):
            block = torch.from_numpy(block)
        if not self.distributed:
            return block.cpu().numpy()
        n = block.shape[0]
        counts = [len(o) for o in owned]
        if block.shape[1] != counts[self.rank] * width:
            raise ValueError("block has {} columns, expected {}".format(block.shape[1], counts[self.rank] * width))
        if block.shape[1] == max(counts) * width:
            padded = block.contiguous()
        else:
            padded = torch.zeros((n, max(counts) * width), dtype=torch.float32, device=block.device)
            padded[:, :block.shape[1]] = block
        full = torch.empty((self.world_size * n, padded.shape[1]), dtype=torch.float32, device=block.device)
        self._dist().all_gather_into_tensor(full, padded)        # rank r's block lands in rows [r*n, (r+1)*n)
        out = np.empty((n, sum(counts) * width), dtype=np.float32)
        host = full.cpu().numpy().reshape(self.world_size, n, padded.shape[1])
        for r, subnets in enumerate(owned):
            for k, s in enumerate(subnets):
                out[:, s * width:(s + 1) * width] = host[r, :, k * width:(k + 1) * width]
        return out

    def gather_blocks_device(self, block, owned, width):
        """``gather_blocks`` that stays on the GPU: returns a float32 CUDA tensor ``[N, S * width]`` with the columns
        in global sub-network order (input of the fused imputation tail, ``Engine.impute(pred=...)``)."""
        import torch
        if not self.distributed:
            return block
        if isinstance(block, np.ndarray):          # gloo runs of the test-suite hand in host arrays
            block = torch.from_numpy(block)
        n = block.shape[0]
        counts = [len(o) for o in owned]
        wmax = max(counts) * width
        if block.shape[1] != counts[self.rank] * width:
            raise ValueError("block has {} columns, expected {}".format(block.shape[1], counts[self.rank] * width))
        if block.shape[1] == wmax:
            padded = block.contiguous()
        else:
            padded = torch.zeros((n, wmax), dtype=torch.float32, device=block.device)
            padded[:, :block.shape[1]] = block
        full = torch.empty((self.world_size * n, wmax), dtype=torch.float32, device=block.device)
        self._dist().all_gather_into_tensor(full, padded)
        out = torch.empty((n, sum(counts) * width), dtype=torch.float32, device=block.device)
        parts = full.view(self.world_size, n, wmax)
        for r, subnets in enumerate(owned):
            for k, s in enumerate(subnets):
                out[:, s * width:(s + 1) * width] = parts[r, :, k * width:(k + 1) * width]
        return out

    def barrier(self):
        if self.distributed:
            self._dist().barrier()


def init(backend=None):
    """Join the process group described by the torchrun environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    have_gpu = torch.cuda.is_available()
    if backend is None:
        backend = "nccl" if have_gpu else "gloo"
    device = local if (have_gpu and backend == "nccl") else None
    if device is not None:
        torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {}
        if device is not None:
            kwargs["device_id"] = torch.device("cuda", device)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return ShardContext(rank, world, device)
