#!/usr/bin/env python
"""Benchmark of the DeepImpute hot path: cells x genes imputed per second over fit + predict.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3|c2|tiny]

One "step" = one pass of the hot path over the workload: ``--epochs`` training epochs of all sub-networks (every
epoch = ceil(n_train / batch) Adam steps + the validation pass) followed by the inference forward over all cells
(reference multinet.py:238-244 + :278-280).  The metric is N_cells * N_genes / (t_fit + t_predict).

  value   device-resident: the normalised matrix is already in HBM, predictions land in an HBM buffer; timed with
          CUDA events on the engine's stream, max over ranks.
  e2e     through the C-ABI with HOST buffers: upload of the normalised matrix from pinned host memory, staging,
          the same epochs, predictions copied back to pinned host memory -- all inside the timed region.
  roofline     the kernel that takes the largest share of a training epoch, timed per launch with CUDA events.
  cpu_baseline the oracle restatement (torch-CPU fp32, one matmul per layer per branch like Keras) on the host
               cores, on a bounded sample, extrapolated to the same workload.

Multi-GPU (torchrun, one rank per GPU): the sub-networks are sharded over the ranks; per epoch the two loss
scalars are all-reduced, per predict the column blocks are all-gathered (NCCL).  Total work is fixed => "strong".

``--impl reference`` times the CPU restatement of the reference path instead (TensorFlow/Keras cannot be
installed offline); only rank 0 works.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2]: the configuration the north-star target is quoted on
    "c3": dict(n_cells=50_000, n_genes=20_000, batch=64, desc="synthetic 50k cells x 20k genes, 40 sub-nets (configs[2])"),
    "c2": dict(n_cells=10_000, n_genes=5_000, batch=64, desc="synthetic 10k cells x 5k genes, 10 sub-nets (configs[1])"),
    # BASELINE.json configs[4]: float32 counts (a float64 frame of this size is 48 GB), batch 256, predictor candidates
    # capped at the 2048 genes with the largest std/mean (the reference's n_pred, multinet.py:25-29)
    "c5": dict(n_cells=200_000, n_genes=30_000, batch=256, n_pred=2048,
               desc="synthetic 200k cells x 30k genes, 59 sub-nets, batch 256, n_pred 2048 (configs[4])"),
    "tiny": dict(n_cells=2_000, n_genes=1_500, batch=64, desc="synthetic 2k cells x 1.5k genes (dry run)"),
}
HIDDEN, OUT, LR, RATE, MODEL_SEED, DATA_SEED = 256, 512, 1e-4, 0.2, 1234, 0


# --------------------------------------------------------------------------------------------- workload
def build_workload(name, device, world=1, rank=0, emulate=None):
    """Synthetic low-rank overdispersed counts and the reference's partition of them.

    ``world`` / ``rank`` (or ``emulate=(rank, world)`` on one GPU): the sub-networks are assigned to ranks and this
    rank's host matrix is COMPACT -- only the gene columns its own sub-networks read (predictors) or fit (targets), in
    ascending gene order, with the index tables renumbered to match -- so a rank uploads 1/N-th of the matrix, not all
    of it.  With one rank the compact matrix is the whole matrix.


    Everything here is set-up, outside every timed region.  The counts are drawn with torch on ``device``; gene
    selection / target assignment / the cell split call the package's own host functions (same np.random stream
    as the reference); predictor selection = top-5 |Pearson r| per target like ``setPredictors`` but evaluated
    with torch (the host float64 corrcoef of 20k genes takes minutes and is not on the measured path).
    Returns norm (pinned host float32 [N, G]), pred_idx (list of int32 arrays), targ_idx [S, O], train/test rows.
    """
    import torch
    from deepimpute_b200 import partition
    w = WORKLOADS[name]
    N, G, K = w["n_cells"], w["n_genes"], 32
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(DATA_SEED)
    Z = torch._standard_gamma(torch.full((N, K), 2.0, device=dev), generator=gen) * 0.5
    W = torch._standard_gamma(torch.full((G, K), 0.3, device=dev), generator=gen)
    scale = torch.exp(torch.randn(G, device=dev, generator=gen))
    raw = torch.empty((N, G), dtype=torch.float32, device=dev)
    for lo in range(0, N, 4096):
        lam = (Z[lo:lo + 4096] @ W.T) * scale * (4.0 / K)
        raw[lo:lo + 4096] = torch.poisson(lam, generator=gen)
    assert float(raw.max()) >= 10          # inspect_data's raw-count check (multinet.py:55-58)

    mean = raw.mean(0, dtype=torch.float64)
    var = ((raw.double() - mean) ** 2).sum(0) / (N - 1) if N * G <= 2e8 else None
    if var is None:                        # chunked to bound memory
        var = torch.zeros(G, dtype=torch.float64, device=dev)
        for lo in range(0, N, 4096):
            var += ((raw[lo:lo + 4096].double() - mean) ** 2).sum(0)
        var /= (N - 1)
    metric = (var / (1 + mean)).cpu().numpy()
    order = np.argsort(-metric, kind="stable")
    order = order[metric[order] > 0]
    np.random.seed(MODEL_SEED)
    genes = partition.choose_genes(order, metric[order], OUT, 0.5, limit=G)       # NN_lim = G pins S (8d)
    targets = partition.assign_targets(genes, OUT)                                # [S, O] gene positions
    cv = var.sqrt() / mean
    cand = (cv > 0) & torch.isfinite(cv)
    if w.get("n_pred"):                    # candidates capped at the n_pred genes with the largest std/mean (multinet.py:25-29)
        cv = torch.where(cand, cv, torch.zeros_like(cv))
        keep = torch.zeros_like(cand)
        keep[cv.argsort(descending=True, stable=True)[:w["n_pred"]]] = True
        cand &= keep

    centred = raw - mean.float()
    centred /= centred.norm(dim=0).clamp_min(1e-30)
    prev = torch.backends.cuda.matmul.allow_tf32 if dev.type == "cuda" else None
    if dev.type == "cuda":
        torch.backends.cuda.matmul.allow_tf32 = False
    corr = (centred.T @ centred).abs_()
    if dev.type == "cuda":
        torch.backends.cuda.matmul.allow_tf32 = prev
    del centred
    corr[:, ~cand] = -1.0
    pred_idx = []
    import pandas as pd
    for t in targets:
        ti = torch.as_tensor(t, device=dev)
        sub = corr[ti].clone()
        sub[:, ti] = -1.0
        top = sub.topk(5, dim=1).indices.reshape(-1).cpu().numpy()
        pred_idx.append(pd.unique(top).astype(np.int32))
    del corr
    from deepimpute_b200 import parallel
    n_pred_all = [len(p) for p in pred_idx]
    if emulate is not None:
        owned = [parallel.assign_subnets(n_pred_all, emulate[1], HIDDEN, OUT)[emulate[0]]]
        mine = owned[0]
    else:
        owned = parallel.assign_subnets(n_pred_all, world, HIDDEN, OUT)
        mine = owned[rank]
    targets = np.ascontiguousarray(targets, dtype=np.int32)
    if world > 1 or emulate is not None:
        cols = np.unique(np.concatenate([pred_idx[s_] for s_ in mine] + [targets[mine].reshape(-1)])).astype(np.int64)
    else:
        cols = np.arange(G, dtype=np.int64)
    renum = np.full(G, -1, dtype=np.int64)
    renum[cols] = np.arange(len(cols))
    norm = torch.empty((N, len(cols)), dtype=torch.float32, pin_memory=(dev.type == "cuda"))
    cols_dev = torch.as_tensor(cols, device=dev)
    for lo in range(0, N, 16384):
        blk = raw[lo:lo + 16384]
        norm[lo:lo + 16384].copy_(torch.log1p(blk if len(cols) == G else blk.index_select(1, cols_dev)))
    raw_keep = raw
    del Z, W
    if dev.type == "cuda":
        torch.cuda.empty_cache()
    np.random.seed(MODEL_SEED)
    train_rows, test_rows = partition.split_cells(N)
    raw_host = None
    # the raw counts are only needed by the single-GPU side blocks (predictor selection, fused post-processing)
    if os.environ.get("DI_BENCH_PREDICTORS", "1") != "0" and dev.type == "cuda" and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        raw_host = torch.empty((N, G), dtype=torch.float32, pin_memory=True)
        raw_host.copy_(raw_keep)
    del raw_keep, raw
    cand_np = torch.nonzero(cand).reshape(-1).cpu().numpy().astype(np.int32)
    return dict(name=name, N=N, G=G, B=w["batch"], norm=norm, raw=raw_host, cand=cand_np,
                # index tables in the coordinates of ``norm``: the whole partition when norm is the whole matrix ...
                pred_idx=pred_idx if len(cols) == G else None, targ_idx=targets if len(cols) == G else None,
                # ... and this rank's sub-networks always
                owned=owned, mine=mine, n_pred_all=n_pred_all, host_cols=len(cols),
                pred_idx_mine=[renum[pred_idx[s_]].astype(np.int32) for s_ in mine],
                targ_idx_mine=np.ascontiguousarray(renum[targets[mine]], dtype=np.int32),
                train_rows=train_rows, test_rows=test_rows, desc=w["desc"])


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) == 6:
                self.samples.append(parts)

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


# --------------------------------------------------------------------------------------- algorithmic work
def kernel_work(n_pred, B):
    """Algorithmic bytes / flops of ONE launch of every training kernel for sub-networks with ``n_pred`` predictors
    (DESIGN.md "Kernels"): fp32 operands; weights + both Adam moments read and written once (24 B / parameter)."""
    S, sumP, H, O = len(n_pred), int(sum(n_pred)), HIDDEN, OUT
    return {
        "fwd1": dict(bytes=4 * (B * sumP + sumP * H + S * H + B * S * H), flops=2 * B * sumP * H),
        "fwd2": dict(bytes=4 * (B * S * H + S * H * O + S * O + 2 * B * S * O), flops=2 * B * S * H * O),
        "bwd": dict(bytes=4 * (B * S * O + S * H * O + 2 * B * S * H), flops=2 * B * S * H * O),
        "adam2": dict(bytes=24 * S * H * O + 4 * B * S * (H + O), flops=2 * B * S * H * O),
        "adam1": dict(bytes=24 * sumP * H + 4 * B * (sumP + S * H), flops=2 * B * sumP * H),
        "bias": dict(bytes=24 * S * (H + O) + 4 * B * S * (H + O), flops=B * S * (H + O)),
        # tensor-core path: both weight matrices in one launch
        "adam": dict(bytes=24 * (sumP * H + S * H * O) + 4 * B * (sumP + 2 * S * H + S * O), flops=2 * B * (sumP * H + S * H * O)),
    }


def step_work(n_pred, B):
    """bytes = sum_s 24 (P H + H + H O + O) + 4 B (P + O); flops = sum_s B (4 P H + 6 H O)."""
    H, O = HIDDEN, OUT
    return (sum(24 * (p * H + H + H * O + O) + 4 * B * (p + O) for p in n_pred),
            sum(B * (4 * p * H + 6 * H * O) for p in n_pred))


def workload_config(wl, epochs):
    """The ``config`` object of the JSON line: what is computed, not how -- both arms print exactly this."""
    n_pred_all = wl.get("n_pred_all") or [len(p) for p in wl["pred_idx"]]
    n_train = len(wl["train_rows"])
    steps_per_epoch = -(-n_train // wl["B"])
    epoch_bytes = step_work(n_pred_all, wl["B"])[0] * steps_per_epoch
    return {"workload": wl["desc"], "epochs_per_step": epochs, "batch_size": wl["B"],
            "sub_networks": len(n_pred_all), "hidden": HIDDEN, "sub_outputdim": OUT,
            "predictors_per_subnet": [int(min(n_pred_all)), int(max(n_pred_all))],
            "adam_steps_per_epoch": steps_per_epoch,
            "l2": "no flush between steps: inputs larger than L2 ({:.1f} GB of weights + Adam state + staged batches "
                  "stream per epoch)".format(epoch_bytes / 1e9)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, tensor=1590.0, source="fallback")


def load_traffic(workload, kernel, n_pred):
    """dram bytes per launch of ``kernel`` from the committed ncu capture (profiles/traffic.json, written by
    scripts/summarise_profile.py from an ``ncu --set full`` run of this same command), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    S, cd = len(n_pred), lambda a, b: -(-a // b)
    max_pp = max(cd(p, 32) * 32 for p in n_pred)
    grids = {"adam": (cd(max_pp, 128) + cd(HIDDEN, 128), max(cd(OUT, 128), cd(HIDDEN, 128)), S),
             "adam2": (cd(HIDDEN, 128), cd(OUT, 128), S), "adam1": (cd(max_pp, 128), cd(HIDDEN, 128), S),
             "fwd1": (1, cd(HIDDEN, 128), S), "fwd2": (1, cd(OUT, 128), S), "bwd": (1, cd(HIDDEN, 128), S)}
    names = {"adam": ("tc_adam_big_kernel", "tc_adam_kernel"), "adam2": "tc_adam_kernel", "adam1": "tc_adam_kernel",
             "fwd1": "tc_kernel<0", "fwd2": "tc_kernel<1", "bwd": "tc_kernel<2"}
    if kernel not in grids:
        return None
    table = json.load(open(path)).get(workload, {})
    if kernel == "adam":
        # the persistent ADAM kernel (the default at many sub-networks per GPU) is launched as a 1-D grid of
        # ceil(tiles / tiles_per_CTA) CTAs
        tiles = S * (cd(max_pp, 128) * cd(HIDDEN, 128) + cd(HIDDEN, 128) * cd(OUT, 128))
        want = "grid ({}, 1, 1)".format(min(148, cd(tiles, int(os.environ.get("DEEPIMPUTE_B200_ADAM_TPC", "2")))))
        for key, val in table.items():
            if key.startswith("tc_adam_pers_kernel") and key.endswith(want):
                return val
    want = "grid ({}, {}, {})".format(*grids[kernel])
    for key, val in table.items():
        if key.startswith(names[kernel]) and key.endswith(want):
            return val
    return None


# ------------------------------------------------------------------------------------------ CPU baseline
def cpu_reference(wl, epochs, n_sub=4, max_epoch_s=None):
    """The oracle restatement timed on the host cores on a bounded sample of the same workload.

    Sample: ``n_sub`` of the S sub-networks (spread over the model), the reference's whole flow for them --
    staging gathers of the train / test matrices (multinet.py:231-235), ONE FULL training epoch (every Adam step of
    the epoch, measured, not extrapolated from a few) plus the validation forward (:238-244), and the inference
    forward over all N cells (:278-280).  The reference runs its branches one after the other inside every step
    (one Keras op dispatch per layer per branch, no cross-branch batching), so cost is linear in the number of
    branches: the sample is scaled by S / n_sub, and the epoch by ``epochs``.  Both factors are in the result.
    ``max_epoch_s``: stop the epoch after that many seconds and scale by the steps done (only for workloads whose
    single epoch would exceed the budget; reported as ``epoch_fraction``).
    """
    import torch
    from oracle.multinet_oracle import OracleNet, stage
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    norm = wl["norm"].numpy() if hasattr(wl["norm"], "numpy") else wl["norm"]
    S = len(wl["pred_idx"])
    n_sub = max(1, min(n_sub, S))
    sel = sorted(set(int(round(i * (S - 1) / max(1, n_sub - 1))) for i in range(n_sub))) if n_sub > 1 else [0]
    pred_idx = [wl["pred_idx"][s] for s in sel]
    targ_idx = wl["targ_idx"][sel]
    n_pred = [len(p) for p in pred_idx]
    B, tr, te = wl["B"], wl["train_rows"], wl["test_rows"]
    net = OracleNet(n_pred, HIDDEN, OUT, learning_rate=LR, batch_size=B, dropout_rate=RATE, seed=MODEL_SEED,
                    mask_mode="torch", subnet_ids=sel)
    t0 = time.perf_counter()
    Xtr, Ytr = stage(norm, pred_idx, targ_idx, tr)
    Xte, Yte = stage(norm, pred_idx, targ_idx, te)
    t_stage = time.perf_counter() - t0
    perm = np.random.default_rng(1).permutation(len(tr)).astype(np.int32)
    steps_per_epoch = -(-len(tr) // B)
    net.train_step([x[:B] for x in Xtr], [y[:B] for y in Ytr], 0)          # warm-up (allocator, thread pool)
    net.set_weights(net.get_weights())
    t0 = time.perf_counter()
    done = 0
    for lo in range(0, len(tr), B):
        rows = perm[lo:lo + B]
        net.train_step([x[rows] for x in Xtr], [y[rows] for y in Ytr], done)
        done += 1
        if max_epoch_s and time.perf_counter() - t0 > max_epoch_s:
            break
    t_steps = time.perf_counter() - t0
    frac = done / float(steps_per_epoch)
    t0 = time.perf_counter()
    net.loss(Xte, Yte)
    t_val = time.perf_counter() - t0
    del Xtr, Ytr
    t0 = time.perf_counter()
    for lo in range(0, wl["N"], 8192):                                     # predict over all cells, chunked for memory
        rows = np.arange(lo, min(lo + 8192, wl["N"]))
        net.forward(stage(norm, pred_idx, targ_idx, rows)[0])
    t_pred_s = time.perf_counter() - t0
    scale = S / float(len(sel))
    t_fit = scale * (t_stage + epochs * (t_steps / frac + t_val))
    t_pred = scale * t_pred_s
    value = wl["N"] * wl["G"] / (t_fit + t_pred)
    sample = ("{} of {} sub-networks {}: staging gathers {:.2f} s, {} of {} Adam steps of one epoch at batch {} in {:.2f} s, "
              "validation forward {:.2f} s, predict over all {} cells {:.2f} s; scaled by S/n = {:.2f} (branches run one "
              "after the other in the reference) and the epoch by {} epochs"
              .format(len(sel), S, sel, t_stage, done, steps_per_epoch, B, t_steps, t_val, wl["N"], t_pred_s, scale, epochs))
    return dict(value=value, unit="cells*genes/s", cores=cores, kind="port", sample=sample,
                sampled_seconds=round(t_stage + t_steps + t_val + t_pred_s, 2),
                extrapolation_factor=round((t_fit + t_pred) / (t_stage + t_steps + t_val + t_pred_s), 1),
                epoch_fraction=round(frac, 4),
                t_fit_s=t_fit, t_predict_s=t_pred)


# ------------------------------------------------------------------------------------------ parity checks
def oracle_check(eng, wl, pred_idx, targ_idx, mine, n_train, steps_cap=None):
    """After the timed regions: re-initialise the very engine that was timed (same handle, same kernels, same
    launch mode), train ONE epoch of the workload and compare one of its sub-networks with the oracle trained alone on
    the same rows (branches share nothing, reference multinet.py:132-148).  Returns the ``parity_check`` object."""
    from deepimpute_b200.engine import epoch_permutation, glorot_uniform
    from oracle.multinet_oracle import OracleNet, stage
    norm = wl["norm"].numpy() if hasattr(wl["norm"], "numpy") else wl["norm"]
    k = len(mine) // 2
    gid = mine[k]
    eng.set_weights(glorot_uniform(eng.n_pred, eng.H, eng.O, MODEL_SEED, mine))
    perm = epoch_permutation(MODEL_SEED, 0, n_train)
    eng.train_epoch(perm, first_step=0)
    rows = np.random.default_rng(7).choice(wl["N"], 1024, replace=False).astype(np.int32)
    got = eng.predict(rows=rows)[:, k * OUT:(k + 1) * OUT]
    got_w = eng.get_weights()[k]
    ref = OracleNet([eng.n_pred[k]], HIDDEN, OUT, learning_rate=LR, batch_size=wl["B"], dropout_rate=RATE,
                    seed=MODEL_SEED, subnet_ids=[gid])
    Xtr, Ytr = stage(norm, [pred_idx[k]], targ_idx[k:k + 1], wl["train_rows"])
    t0 = time.perf_counter()
    ref.train_epoch(Xtr, Ytr, perm, 0)
    secs = time.perf_counter() - t0
    want = ref.forward(stage(norm, [pred_idx[k]], targ_idx[k:k + 1], rows)[0])[0]
    rel = lambda a, b: float(np.max(np.abs(np.asarray(a, np.float64) - b)) / (np.max(np.abs(b)) + 1e-300))   # noqa: E731
    max_rel_w = max(rel(a, b) for a, b in zip(got_w, ref.get_weights()[0]))
    max_rel = rel(got, want)
    tol = 1e-3            # the north star's bound; measured 1e-6 .. 5e-5 (profiles/r02e_accumulation.md)
    return {"what": "sub-network {} of the timed engine after one epoch ({} Adam steps) from its initial weights vs the "
                    "CPU oracle trained alone on the same rows".format(gid, -(-n_train // wl["B"])),
            # ok is decided on the predictions (the north star bounds the imputed values); the weights are reported
            # beside it: after hundreds of steps a few weights of rarely-active hidden units differ between ANY two fp32
            # implementations (the fp32 CUDA-core engine itself: 1e-2 against the oracle after 320 steps at batch 256,
            # profiles/traces/r02u_family_accuracy_batch256.txt)
            "max_rel": max_rel, "max_rel_weights": max_rel_w, "tol": tol, "ok": bool(max_rel < tol),
            "oracle_seconds": round(secs, 1)}


def multinet_shard_check(ctx, rank, world, local):
    """``--gpus N``: the reference-facing object itself, sharded.  ``MultiNet(shard=ctx).fit(df).predict(df)`` on every
    rank (sub-networks split over the ranks, NCCL all-reduce of the epoch losses, NCCL all-gather of the prediction
    blocks, fused imputation tail on every rank) must return the imputed frame of the unsharded ``MultiNet`` bit for bit."""
    import contextlib
    import pandas as pd
    from deepimpute_b200 import MultiNet
    rng = np.random.default_rng(5)
    n_cells, n_genes, rank_k = 640, 2400, 8
    lam = (rng.gamma(2.0, 0.5, size=(n_cells, rank_k)) @ rng.gamma(0.3, 1.0, size=(n_genes, rank_k)).T) * \
        rng.lognormal(0.0, 1.0, size=n_genes) * (4.0 / rank_k)
    counts = rng.poisson(lam).astype(np.float64)
    counts[0, 0] = max(counts[0, 0], 12.0)
    frame = pd.DataFrame(counts, index=["c{}".format(i) for i in range(n_cells)], columns=["g{}".format(j) for j in range(n_genes)])
    kw = dict(seed=1234, ncores=1, max_epochs=3, patience=100, verbose=0, sub_outputdim=128,
              architecture=[{"type": "dense", "neurons": 64, "activation": "relu"}, {"type": "dropout", "rate": 0.2}])
    with contextlib.redirect_stdout(sys.stderr):
        net = MultiNet(shard=ctx, device=local, **kw)
        net.fit(frame, NN_lim=2048)
        out = net.predict(frame)
    ctx.barrier()
    result = None
    if rank == 0:
        try:                                     # rank 0 works alone here: whatever happens, it must reach the barrier below
            with contextlib.redirect_stdout(sys.stderr):
                one = MultiNet(device=local, **kw)
                one.fit(frame, NN_lim=2048)
                ref = one.predict(frame)
            loss_rel = float(np.max(np.abs(np.asarray(net.history["loss"]) - np.asarray(one.history["loss"])) /
                                    np.abs(np.asarray(one.history["loss"]))))
            diff = float(np.max(np.abs(out.values - ref.values)))
            result = {"what": "MultiNet(shard=ctx).fit/predict, {} sub-networks over {} ranks, vs the unsharded MultiNet"
                              .format(len(net.predictors), world),
                      "max_abs_diff_imputed": diff, "max_rel_diff_losses": loss_rel,
                      "ok": bool(diff == 0.0 and loss_rel < 1e-5)}
            if one.engine is not None:
                one.engine.close()
        except Exception as exc:
            result = {"ok": False, "error": repr(exc)}
    if net.engine is not None:
        net.engine.close()
    ctx.barrier()
    return result


def sharding_check(ctx, rank, world, local):
    """``--gpus N``: a small model sharded over the N ranks must give the prediction blocks of the unsharded model
    bit for bit (global sub-network numbers key the initial weights and the dropout stream).  Rank 0 trains both."""
    import torch
    from deepimpute_b200 import parallel
    from deepimpute_b200.engine import Engine, epoch_permutation
    S, O, H, B, N = 2 * world + 1, 64, 48, 32, 700
    G = S * O + 260                              # disjoint targets for every sub-network plus spare genes
    rng = np.random.default_rng(3)
    lam = rng.gamma(0.6, 3.0, size=(1, G)) * rng.gamma(2.0, 0.5, size=(N, 1))
    norm = np.log1p(rng.poisson(lam)).astype(np.float32)
    n_pred = [int(x) for x in rng.integers(40, 90, size=S)]
    perm_g = rng.permutation(G)
    targ = perm_g[:S * O].reshape(S, O).astype(np.int32)
    pred = [rng.choice(G, p, replace=False).astype(np.int32) for p in n_pred]
    tr, te = np.arange(0, 640, dtype=np.int32), np.arange(640, N, dtype=np.int32)
    owned = parallel.assign_subnets(n_pred, world, H, O)

    def run(ids):
        e = Engine([n_pred[s] for s in ids], hidden=H, sub_outputdim=O, batch_size=B, seed=5, device=local, subnet_ids=ids)
        e.set_data(norm, [pred[s] for s in ids], targ[ids])
        e.set_split(tr, te)
        losses = [e.train_epoch(epoch_permutation(5, ep, len(tr))) for ep in range(2)]
        out = e.predict()
        e.close()
        return losses, out

    losses, block = run(owned[rank])
    width = max(len(o) for o in owned) * O
    mine = torch.zeros((N, width), dtype=torch.float32, device="cuda")
    mine[:, :block.shape[1]] = torch.from_numpy(block).cuda()
    allb = torch.empty((world * N, width), dtype=torch.float32, device="cuda")
    torch.distributed.all_gather_into_tensor(allb, mine)
    l = torch.tensor(losses, dtype=torch.float64, device="cuda")
    torch.distributed.all_reduce(l)
    if rank != 0:
        return None
    full_losses, full = run(list(range(S)))          # (no collective follows: rank 0 may take its time, or fail, alone)
    allb = allb.cpu().numpy().reshape(world, N, width)
    diff = 0.0
    for r, ids in enumerate(owned):
        for k, s in enumerate(ids):
            diff = max(diff, float(np.max(np.abs(allb[r][:, k * O:(k + 1) * O] - full[:, s * O:(s + 1) * O]))))
    loss_rel = float(np.max(np.abs(l.cpu().numpy() - np.asarray(full_losses)) / np.abs(np.asarray(full_losses))))
    return {"what": "{} sub-networks sharded over {} ranks vs one engine holding all of them, 2 epochs + predict"
                    .format(S, world),
            "max_abs_diff_predictions": diff, "max_rel_diff_summed_losses": loss_rel,
            "ok": bool(diff == 0.0 and loss_rel < 1e-5)}


# ------------------------------------------------------------------------------------------- API level
def api_run(wl, epochs, local):
    """``MultiNet(...).fit(df, NN_lim=G)`` + ``.predict(df)`` on the workload's raw counts as a float32 DataFrame -- the
    call a user of the reference makes (multinet.py:169, :266), everything included: input checks, gene statistics,
    correlation + predictor selection, upload, the epochs, model file, held-out metrics, predict with the fused tail
    and the float64 DataFrame that comes back.  Twice: ``epochs`` fixed epochs (patience off), and the reference's
    defaults (max_epochs 500, patience 5: early-stopped)."""
    import contextlib
    import pandas as pd
    from deepimpute_b200 import MultiNet
    N, G = wl["N"], wl["G"]
    frame = pd.DataFrame(wl["raw"].numpy(), index=["c{:06d}".format(i) for i in range(N)],
                         columns=["g{:06d}".format(j) for j in range(G)], copy=False)
    out = {}
    for name, kw in (("fixed_epochs", dict(max_epochs=epochs, patience=10 ** 6)), ("early_stopped", dict(max_epochs=500, patience=5))):
        net = MultiNet(seed=MODEL_SEED, ncores=1, verbose=0, device=local, **kw)
        with contextlib.redirect_stdout(sys.stderr):
            t0 = time.perf_counter()
            net.fit(frame, NN_lim=G)
            t1 = time.perf_counter()
            imputed = net.predict(frame)
            t2 = time.perf_counter()
        assert imputed.shape == (N, G)
        out[name] = {"fit_s": round(t1 - t0, 3), "predict_s": round(t2 - t1, 3), "epochs": int(net.trained_epochs),
                     "value": N * G / (t2 - t0), "unit": "cells*genes/s",
                     "sub_networks": int(len(net.predictors)),
                     "test_metrics": {k: float(v) for k, v in net.test_metrics.items()},
                     "phases": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in net.timings.items()}}
        if net.engine is not None:
            net.engine.close()
        del net, imputed
    return out


# --------------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("DI_BENCH_WORKLOAD", "c3"), choices=sorted(WORKLOADS))
    ap.add_argument("--epochs", type=int, default=20, help="training epochs per step (fixed; no early stopping)")
    ap.add_argument("--math", default=None, choices=["fp32", "tf32", "tf32x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--api", action="store_true",
                    help="also time the reference-facing API end to end: MultiNet(...).fit(DataFrame, NN_lim=G) + "
                         ".predict(DataFrame) on the workload's count matrix, fixed epochs and early-stopped (N = 1 only)")
    ap.add_argument("--ref-subnets", type=int, default=4, help="sub-networks in the CPU sample of --impl reference")
    ap.add_argument("--no-checks", action="store_true", help="skip the oracle / sharding checks after the timed regions")
    ap.add_argument("--emulate-shard", default=None, metavar="R/N",
                    help="tuning aid: on ONE GPU, train and predict only the sub-networks rank R of an N-rank run would own "
                         "(what one GPU of an N-GPU job does, without the collectives); the line says so in config")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "cells x genes imputed/sec (fit+predict)"
    unit = "cells*genes/s"

    import torch

    # ------------------------------------------------------------------------ reference arm (CPU restatement)
    if args.impl == "reference":
        if rank != 0:
            return 0
        dev = "cuda:0" if torch.cuda.is_available() else "cpu"
        wl = build_workload(args.workload, dev)
        for _ in range(args.warmup):                   # thread pool, allocator, page cache: a short slice of the same sample
            cpu_reference(wl, args.epochs, n_sub=1, max_epoch_s=1.0)
        runs = [cpu_reference(wl, args.epochs, n_sub=args.ref_subnets) for _ in range(max(1, args.steps))]
        value = float(np.mean([r["value"] for r in runs]))
        secs = wl["N"] * wl["G"] / value
        base = dict(runs[-1], value=value)
        for k in ("t_fit_s", "t_predict_s"):
            base.pop(k, None)
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(wl, args.epochs),
            "arm": {"note": "restated reference (TensorFlow/Keras unavailable offline): torch-CPU fp32, one "
                            "matmul per layer per branch.  Each step MEASURES the staging gathers, one full epoch, the "
                            "validation forward and the predict pass of cpu_baseline.sample's sub-networks and scales "
                            "by the stated factor; ms_per_step is that extrapolated fit+predict time",
                    "sampled_seconds_per_step": base.get("sampled_seconds"),
                    "extrapolation_factor": base.get("extrapolation_factor")},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }))
        return 0

    # ------------------------------------------------------------------------------------------- B200 arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference)")
    from deepimpute_b200 import parallel
    from deepimpute_b200.engine import DEFAULT_MATH, Engine, PermutationPrefetcher, epoch_permutation
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout to the one JSON line (NCCL's version banner)
    ctx = parallel.init()
    torch.cuda.set_device(local)
    emulate = tuple(int(x) for x in args.emulate_shard.split("/")) if (args.emulate_shard and world == 1) else None
    wl = build_workload(args.workload, "cuda:{}".format(local), world, rank, emulate)
    N, G, B = wl["N"], wl["G"], wl["B"]
    n_pred_all, owned, mine = wl["n_pred_all"], wl["owned"], wl["mine"]
    S_all = len(n_pred_all)
    n_pred = [n_pred_all[s] for s in mine]
    pred_idx, targ_idx = wl["pred_idx_mine"], wl["targ_idx_mine"]       # columns of this rank's (compact) host matrix
    math_mode = args.math or os.environ.get("DEEPIMPUTE_B200_MATH", DEFAULT_MATH)
    eng = Engine(n_pred, hidden=HIDDEN, sub_outputdim=OUT, learning_rate=LR, batch_size=B, dropout_rate=RATE,
                 seed=MODEL_SEED, math_mode=math_mode, device=local, subnet_ids=mine)
    norm_np = wl["norm"].numpy()
    n_train = len(wl["train_rows"])
    steps_per_epoch = -(-n_train // B)
    state = {"epoch": 0}
    width = len(mine) * OUT
    pad_width = max(len(o) for o in owned) * OUT
    out_dev = torch.empty((N, pad_width), dtype=torch.float32, device="cuda")
    gathered = torch.empty((world * N, pad_width), dtype=torch.float32, device="cuda") if world > 1 else None

    # the visiting order of the next epoch is drawn on a helper thread while the device runs the current one, as
    # Engine.fit does (about 1 ms of numpy per epoch that the GPU would otherwise wait for)
    perms = PermutationPrefetcher(lambda e: epoch_permutation(MODEL_SEED, e, n_train))

    def train_epochs():
        for _ in range(args.epochs):
            loss, val = eng.train_epoch(perms.get(state["epoch"]))
            state["epoch"] += 1
            if world > 1:                      # EarlyStopping watches the sum over all branches (multinet.py:242)
                loss, val = ctx.sum_scalars(loss, val)
        return loss, val

    def device_step():
        train_epochs()
        eng.predict_device(out_dev.data_ptr(), pad_width)
        if world > 1:
            torch.distributed.all_gather_into_tensor(gathered, out_dev)
            torch.cuda.synchronize()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        eng.sync()

    eng.set_data(norm_np, pred_idx, targ_idx)
    eng.set_split(wl["train_rows"], wl["test_rows"])
    for _ in range(args.warmup):
        device_step()
    barrier()
    launches0 = eng.launch_count()
    with ClockSampler(local) as clocks:
        eng.timer_start()
        for _ in range(args.steps):
            device_step()
        torch.cuda.synchronize()
        dev_ms = eng.timer_stop()
    barrier()
    launches = eng.launch_count() - launches0

    # ---- end to end through the C-ABI with host buffers
    # every rank owns the host copy of its own block of imputed columns: nothing is funnelled through rank 0
    host_block = torch.empty((N, width if world == 1 else pad_width), dtype=torch.float32, pin_memory=True)
    host_out = host_block.numpy()

    def e2e_step():
        eng.set_data(norm_np, pred_idx, targ_idx)                   # H2D of this rank's (compact) normalised matrix
        eng.set_split(wl["train_rows"], wl["test_rows"])
        loss, _ = train_epochs()
        if world == 1:
            eng.predict(out=host_out)                               # D2H of the imputed block
        else:
            eng.predict_device(out_dev.data_ptr(), pad_width)
            torch.distributed.all_gather_into_tensor(gathered, out_dev)      # reassembled on every GPU over NVLink
            host_block.copy_(out_dev, non_blocking=True)                     # D2H of this rank's own block
            torch.cuda.synchronize()
        return loss

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    perms.close()

    # ---- per-kernel timing of one more epoch (CUDA events around every launch on the engine's stream)
    eng.set_profiling(True)
    eng.train_epoch(epoch_permutation(MODEL_SEED, state["epoch"], n_train))
    names = ["gather", "fwd1", "fwd2", "bwd", "adam", "adam2", "adam1", "bias", "infer1", "infer2"]
    kern = {k: (eng.kernel_ms(k), eng.kernel_launches(k)) for k in names if eng.kernel_launches(k) > 0}
    eng.set_profiling(False)
    epoch_ms = eng.last_device_ms()
    # the inference pass on its own (the tensor-pipe bound part of the path): device time of one predict
    predict_ms = None
    try:
        eng.predict_device(out_dev.data_ptr(), pad_width)
        predict_ms = float(eng.last_device_ms())
    except Exception as exc:                                            # never lose the bench line over a side figure
        print("bench.py: predict timing skipped ({})".format(exc), file=sys.stderr)

    # max over ranks
    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    dev_ms, e2e_s = float(t[0]), float(t[1])
    h2d = int(N * wl["host_cols"] * 4 + args.epochs * n_train * 4)
    d2h = int(N * (width if world == 1 else pad_width) * 4)
    if world > 1:
        tl = torch.tensor([launches, h2d, d2h], dtype=torch.int64, device="cuda")
        torch.distributed.all_reduce(tl)
        launches, h2d, d2h = int(tl[0]), int(tl[1]), int(tl[2])

    # ---- correctness of what was timed (outside every timed region)
    checks = {}
    if not args.no_checks:
        if world > 1:
            for name, fn in (("sharding_check", sharding_check), ("multinet_shard_check", multinet_shard_check)):
                try:
                    checks[name] = fn(ctx, rank, world, local)
                except Exception as exc:                                # report, never lose the bench line (every rank fails alike)
                    checks[name] = {"ok": False, "error": repr(exc)}
        elif wl["name"] != "c5" or os.environ.get("DI_BENCH_ORACLE_C5") == "1":
            # one oracle epoch of one sub-network: ~5 s at c3 (batch 64); ~1 min at c5 (batch 256), on request only
            try:
                eng.set_data(norm_np, pred_idx, targ_idx)
                eng.set_split(wl["train_rows"], wl["test_rows"])
                checks["parity_check"] = oracle_check(eng, wl, pred_idx, targ_idx, mine, n_train)
            except Exception as exc:                                    # report, never lose the bench line
                checks["parity_check"] = {"ok": False, "error": repr(exc)}

    if rank == 0:
        ms_per_step = dev_ms / args.steps
        value = N * G / (ms_per_step * 1e-3)
        peaks = load_peaks()
        work = kernel_work(n_pred, B)
        per_kernel = {}
        for k, (ms, cnt) in kern.items():
            entry = {"ms": round(ms, 5), "launches": cnt, "share_of_epoch": round(ms * cnt / epoch_ms, 4)}
            if k in work:
                entry["GB/s"] = round(work[k]["bytes"] / (ms * 1e-3) / 1e9, 1)
                entry["TFLOP/s"] = round(work[k]["flops"] / (ms * 1e-3) / 1e12, 2)
            per_kernel[k] = entry
        train_kernels = [k for k in per_kernel if k in work]
        top = max(train_kernels, key=lambda k: per_kernel[k]["ms"] * per_kernel[k]["launches"])
        achieved = work[top]["bytes"] / (kern[top][0] * 1e-3) / 1e9
        step_bytes, step_flops = step_work(n_pred, B)
        step_ms = sum(per_kernel[k]["ms"] for k in train_kernels)
        roofline = {"bound": "hbm", "kernel": top, "achieved": round(achieved, 1), "peak": peaks["hbm"],
                    "unit": "GB/s", "frac": round(achieved / peaks["hbm"], 4),
                    "traffic": load_traffic(wl["name"], top, n_pred), "peak_source": peaks["source"],
                    "algorithmic_bytes_per_launch": work[top]["bytes"],
                    "predict": None if not predict_ms else {
                        "ms": round(predict_ms, 3), "bound": "tensor",
                        "TFLOP/s_fp32_equivalent": round(sum(2.0 * N * (p * HIDDEN + HIDDEN * OUT) for p in n_pred)
                                                         / (predict_ms * 1e-3) / 1e12, 2),
                        "note": "gathers + FWD1 + FWD2 over all cells of this rank's sub-networks; tf32x3 issues three "
                                "TF32 products per fp32-equivalent product"},
                    "train_step": {"ms": round(step_ms, 4), "GB/s": round(step_bytes / (step_ms * 1e-3) / 1e9, 1),
                                   "frac": round(step_bytes / (step_ms * 1e-3) / 1e9 / peaks["hbm"], 4),
                                   "TFLOP/s": round(step_flops / (step_ms * 1e-3) / 1e12, 2),
                                   "algorithmic_bytes": step_bytes,
                                   "note": "sum of the four kernels launched one by one (profiling epoch), not overlapped"},
                    # what the driver's clock sees: the timed region divided by its Adam steps (the per-epoch staging
                    # gather and validation pass included), against the same algorithmic bytes
                    "train_step_timed": (lambda us: {
                        "us": round(us, 2), "GB/s": round(step_bytes / (us * 1e-6) / 1e9, 1),
                        "frac": round(step_bytes / (us * 1e-6) / 1e9 / peaks["hbm"], 4)})(
                        (ms_per_step - (predict_ms or 0.0)) * 1e3 / (args.epochs * steps_per_epoch)),
                    "kernels": per_kernel}
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32", "tf32x3": "tf32x3 (error-compensated TF32, fp32 accumulate)"}[math_mode], "data": "synthetic",
            "config": workload_config(wl, args.epochs),
            "arm": {"math_mode": math_mode,
                    "parallelism": "sub-networks sharded over {} GPU(s)".format(world) if not args.emulate_shard else
                                   "EMULATION of rank {} of a sharded run on one GPU: {} of {} sub-networks, no collectives"
                                   .format(args.emulate_shard, len(mine), S_all),
                    "epoch_driver": "one CUDA-graph launch per epoch over sub-network groups on concurrent streams; "
                                    "roofline.kernels are timed in a separate launch-by-launch epoch"},
            "clocks": clocks.summary(),
            "e2e": {"value": N * G / e2e_s, "unit": unit, "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "summed over ranks: every rank uploads the gene columns its own sub-networks use and copies "
                            "its own block of imputed columns back to its own pinned buffer"},
            "gpu_launches": launches,
            "engine": eng.describe(), "graph_fallbacks": eng.graph_fallbacks(),
            "roofline": roofline,
        }
        line.update({k: v for k, v in checks.items() if v is not None})
        if world == 1 and wl.get("raw") is not None:
            # SURVEY.md 8f row 1, outside every timed region above: correlation + top-5 predictor selection of the
            # same matrix through di_corr_topk (host buffers in, indices out), checked against the set-up's selection
            from deepimpute_b200 import partition as _part
            t0 = time.perf_counter()
            labels = np.array(["g{:06d}".format(j) for j in wl["cand"]], dtype=object)
            import contextlib
            with contextlib.redirect_stdout(sys.stderr):          # the helper prints one line per sub-network
                picked, dev_ms = _part.choose_predictors_gpu(wl["raw"].numpy(), wl["targ_idx"], wl["cand"], labels, 5,
                                                             device=local)
            secs = time.perf_counter() - t0
            same = sum(len(np.intersect1d(a, b)) for a, b in zip(picked, wl["pred_idx"])) / float(sum(n_pred_all))
            line["predictor_selection"] = {
                "seconds": round(secs, 3), "device_ms": round(dev_ms, 1),
                "TFLOP/s": round(2.0 * G * G * N / (dev_ms * 1e-3) / 1e12, 2),
                "overlap_with_setup_selection": round(same, 5),
                "note": "|corrcoef| of raw counts + per-target top-5 (multinet.py:20-34, :344-365); fp32 CUDA-core Gram kernel"}
        if world == 1 and wl.get("raw") is not None and os.environ.get("DI_BENCH_IMPUTE", "1") != "0":
            # SURVEY.md 8f rows 2+3, outside every timed region above: raw counts up, log1p on the device, then the fused
            # tail of MultiNet.predict (forward + duplicate mean + clamp + expm1 + restore) into a float64 [N, G] host matrix
            raw_np = wl["raw"].numpy()
            imputed = torch.empty((N, G), dtype=torch.float64, pin_memory=True).numpy()
            eng.set_profiling(True)
            t0 = time.perf_counter()
            eng.set_counts(raw_np, pred_idx, targ_idx)
            t1 = time.perf_counter()
            eng.impute(policy="restore", out=imputed)
            t2 = time.perf_counter()
            log1p_ms, impute_ms, n_imp = eng.kernel_ms("log1p"), eng.kernel_ms("impute"), eng.kernel_launches("impute")
            eng.set_profiling(False)
            kept = bool(np.array_equal(imputed[:64][raw_np[:64] > 0], raw_np[:64][raw_np[:64] > 0].astype(np.float64)))
            line["postprocess"] = {
                "upload_counts_s": round(t1 - t0, 4), "impute_s": round(t2 - t1, 4),
                "log1p_kernel": {"ms": round(log1p_ms, 4), "GB/s": round(8.0 * N * G / (log1p_ms * 1e-3) / 1e9, 1),
                                 "algorithmic_bytes": 8 * N * G},
                "impute_kernel": {"ms_per_chunk": round(impute_ms, 4), "chunks": n_imp,
                                  "GB/s": round((12.0 * N * G + 4.0 * N * S_all * OUT) / (impute_ms * n_imp * 1e-3) / 1e9, 1),
                                  "algorithmic_bytes": int(12 * N * G + 4 * N * S_all * OUT)},
                "d2h_bytes": int(8 * N * G), "d2h_GB/s_incl_forward": round(8.0 * N * G / (t2 - t1) / 1e9, 1),
                "restore_keeps_observed_counts": kept,
                "note": "di_upload_counts + di_impute (multinet.py:217/:271 and :278-303); float64 [N, G] out like the "
                        "reference's DataFrame; the host-side pandas route needs several N x G float64 temporaries"}
            del imputed
        if world == 1 and args.api and wl.get("raw") is not None:
            eng.close()                                    # the API run builds its own engines
            line["api"] = api_run(wl, args.epochs, local)
        if world == 1 and not args.no_cpu_baseline and wl["pred_idx"] is not None:
            # bounded: 2 sub-networks, at most ~15 s of Adam steps (the reference arm times 4 and the whole epoch)
            line["cpu_baseline"] = {k: v for k, v in cpu_reference(wl, args.epochs, n_sub=2, max_epoch_s=15.0).items()
                                    if k not in ("t_fit_s", "t_predict_s")}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
