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
    "tiny": dict(n_cells=2_000, n_genes=1_500, batch=64, desc="synthetic 2k cells x 1.5k genes (dry run)"),
}
HIDDEN, OUT, LR, RATE, MODEL_SEED, DATA_SEED = 256, 512, 1e-4, 0.2, 1234, 0


# --------------------------------------------------------------------------------------------- workload
def build_workload(name, device):
    """Synthetic low-rank overdispersed counts and the reference's partition of them.

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
    cand = ((var.sqrt() / mean) > 0) & torch.isfinite(var.sqrt() / mean)

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
    norm = torch.empty((N, G), dtype=torch.float32, pin_memory=(dev.type == "cuda"))
    norm.copy_(torch.log1p(raw))
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
    return dict(name=name, N=N, G=G, B=w["batch"], norm=norm, pred_idx=pred_idx, raw=raw_host, cand=cand_np,
                targ_idx=np.ascontiguousarray(targets, dtype=np.int32), train_rows=train_rows, test_rows=test_rows,
                desc=w["desc"])


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
    want = "grid ({}, {}, {})".format(*grids[kernel])
    for key, val in json.load(open(path)).get(workload, {}).items():
        if key.startswith(names[kernel]) and key.endswith(want):
            return val
    return None


# ------------------------------------------------------------------------------------------ CPU baseline
def cpu_reference(wl, epochs, budget_s=12.0, min_steps=2):
    """The oracle restatement timed on the host cores on a bounded sample, extrapolated to the workload.

    Sample: ``n`` optimiser steps (all S sub-networks, batch B) and one inference forward over 1024 cells.
    t_fit = epochs * (steps_per_epoch * t_step + n_test/1024 * t_fwd);  t_predict = N/1024 * t_fwd.
    """
    import torch
    from oracle.multinet_oracle import OracleNet, stage
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    norm = wl["norm"].numpy() if hasattr(wl["norm"], "numpy") else wl["norm"]
    n_pred = [len(p) for p in wl["pred_idx"]]
    B = wl["B"]
    net = OracleNet(n_pred, HIDDEN, OUT, learning_rate=LR, batch_size=B, dropout_rate=RATE, seed=MODEL_SEED,
                    mask_mode="torch")
    rng = np.random.default_rng(1)

    def batch():
        rows = np.sort(rng.choice(wl["train_rows"], B, replace=False))
        return stage(norm, wl["pred_idx"], wl["targ_idx"], rows)

    X, Y = batch()
    net.train_step(X, Y, 0)                                    # warm-up (allocator, thread pool)
    n, t_steps = 0, 0.0
    while n < min_steps or (t_steps < budget_s * 0.7 and n < 200):
        X, Y = batch()
        t0 = time.perf_counter()
        net.train_step(X, Y, n + 1)
        t_steps += time.perf_counter() - t0
        n += 1
    t_step = t_steps / n
    rows = np.arange(min(1024, wl["N"]))
    Xf, _ = stage(norm, wl["pred_idx"], wl["targ_idx"], rows)
    net.forward(Xf)
    reps, t_fwd = 0, 0.0
    while reps < 1 or (t_fwd < budget_s * 0.3 and reps < 20):
        t0 = time.perf_counter()
        net.forward(Xf)
        t_fwd += time.perf_counter() - t0
        reps += 1
    t_fwd = t_fwd / reps * (1024.0 / len(rows))
    steps_per_epoch = -(-len(wl["train_rows"]) // B)
    t_fit = epochs * (steps_per_epoch * t_step + len(wl["test_rows"]) / 1024.0 * t_fwd)
    t_pred = wl["N"] / 1024.0 * t_fwd
    value = wl["N"] * wl["G"] / (t_fit + t_pred)
    sample = ("{} Adam steps of all {} sub-networks at batch {} ({:.3f} s/step) + forward over {} cells "
              "({:.3f} s per 1024); extrapolated to {} epochs x {} steps + predict over {} cells"
              .format(n, len(n_pred), B, t_step, len(rows), t_fwd, epochs, steps_per_epoch, wl["N"]))
    return dict(value=value, unit="cells*genes/s", cores=cores, kind="port", sample=sample,
                t_fit_s=t_fit, t_predict_s=t_pred, sampled_s=t_steps + t_fwd * reps)


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
        for _ in range(args.warmup):
            cpu_reference(wl, args.epochs, budget_s=2.0)
        runs = [cpu_reference(wl, args.epochs, budget_s=12.0) for _ in range(max(1, args.steps))]
        value = float(np.mean([r["value"] for r in runs]))
        secs = wl["N"] * wl["G"] / value
        base = dict(runs[-1], value=value)
        for k in ("t_fit_s", "t_predict_s", "sampled_s"):
            base.pop(k, None)
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "epochs_per_step": args.epochs, "batch_size": wl["B"],
                       "sub_networks": len(wl["pred_idx"]), "hidden": HIDDEN, "sub_outputdim": OUT,
                       "note": "restated reference (TensorFlow/Keras unavailable offline): torch-CPU fp32, one "
                               "matmul per layer per branch; each step is a bounded sample extrapolated"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }))
        return 0

    # ------------------------------------------------------------------------------------------- B200 arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference)")
    from deepimpute_b200 import parallel
    from deepimpute_b200.engine import DEFAULT_MATH, Engine, epoch_permutation
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout to the one JSON line (NCCL's version banner)
    ctx = parallel.init()
    torch.cuda.set_device(local)
    wl = build_workload(args.workload, "cuda:{}".format(local))
    N, G, B = wl["N"], wl["G"], wl["B"]
    n_pred_all = [len(p) for p in wl["pred_idx"]]
    S_all = len(n_pred_all)
    owned = parallel.assign_subnets(n_pred_all, world, HIDDEN, OUT)
    mine = owned[rank]
    if args.emulate_shard and world == 1:
        r, n = (int(x) for x in args.emulate_shard.split("/"))
        owned = [parallel.assign_subnets(n_pred_all, n, HIDDEN, OUT)[r]]
        mine = owned[0]
    n_pred = [n_pred_all[s] for s in mine]
    pred_idx = [wl["pred_idx"][s] for s in mine]
    targ_idx = np.ascontiguousarray(wl["targ_idx"][mine])
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

    def train_epochs():
        for _ in range(args.epochs):
            loss, val = eng.train_epoch(epoch_permutation(MODEL_SEED, state["epoch"], n_train))
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
    host_out = torch.empty((N, width), dtype=torch.float32, pin_memory=True).numpy()
    full_host = torch.empty((N, pad_width * world), dtype=torch.float32, pin_memory=True) if (world > 1 and rank == 0) else None

    def e2e_step():
        eng.set_data(norm_np, pred_idx, targ_idx)                   # H2D of the whole normalised matrix
        eng.set_split(wl["train_rows"], wl["test_rows"])
        loss, _ = train_epochs()
        if world == 1:
            eng.predict(out=host_out)                               # D2H of the imputed block
        else:
            eng.predict_device(out_dev.data_ptr(), pad_width)
            torch.distributed.all_gather_into_tensor(gathered, out_dev)
            if rank == 0:
                full_host.view(world * N, pad_width).copy_(gathered, non_blocking=True)
            torch.cuda.synchronize()
        return loss

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps

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
    if world > 1:
        tl = torch.tensor([launches], dtype=torch.int64, device="cuda")
        torch.distributed.all_reduce(tl)
        launches = int(tl[0])

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
                                   "algorithmic_bytes": step_bytes},
                    "kernels": per_kernel}
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32", "tf32x3": "tf32x3 (error-compensated TF32, fp32 accumulate)"}[math_mode], "data": "synthetic",
            "config": {"workload": wl["desc"], "epochs_per_step": args.epochs, "batch_size": B,
                       "sub_networks": S_all, "hidden": HIDDEN, "sub_outputdim": OUT,
                       "predictors_per_subnet": [int(min(n_pred_all)), int(max(n_pred_all))],
                       "adam_steps_per_epoch": steps_per_epoch, "math_mode": math_mode,
                       "parallelism": "sub-networks sharded over {} GPU(s)".format(world) if not args.emulate_shard else
                                      "EMULATION of rank {} of a sharded run on one GPU: {} of {} sub-networks, no collectives"
                                      .format(args.emulate_shard, len(mine), S_all),
                       "epoch_driver": "one CUDA-graph launch per epoch over sub-network groups on concurrent streams; "
                                       "roofline.kernels are timed in a separate launch-by-launch epoch",
                       "l2": "inputs larger than L2: {:.1f} GB of weights+Adam state+staged batches stream per epoch"
                             .format((step_bytes * steps_per_epoch) / 1e9)},
            "clocks": clocks.summary(),
            "e2e": {"value": N * G / e2e_s, "unit": unit, "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": int(N * G * 4 + args.epochs * n_train * 4),
                    "d2h_bytes_per_step": int(N * (width if world == 1 else pad_width * world) * 4)},
            "gpu_launches": launches,
            "engine": eng.describe(), "graph_fallbacks": eng.graph_fallbacks(),
            "roofline": roofline,
        }
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
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = {k: v for k, v in cpu_reference(wl, args.epochs).items()
                                    if k not in ("t_fit_s", "t_predict_s", "sampled_s")}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
