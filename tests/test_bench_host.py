"""Host-side pieces of bench.py that run without a GPU: the per-rank compact matrix of a sharded run must hold exactly
the values the full matrix holds at the renumbered positions, and the CPU arm must produce a consistent line."""
import json
import subprocess
import sys

import numpy as np

import bench
from conftest import ROOT


def test_compact_per_rank_matrix_matches_the_full_one():
    full = bench.build_workload("tiny", "cpu")
    assert full["host_cols"] == full["G"] and full["pred_idx"] is not None
    norm = full["norm"].numpy()
    world = 3
    seen = set()
    for rank in range(world):
        part = bench.build_workload("tiny", "cpu", world, rank)
        mine = part["mine"]
        assert part["owned"] == full["owned"] or len(part["owned"]) == world
        seen.update(mine)
        assert part["pred_idx"] is None and part["host_cols"] < full["G"]        # a rank holds only its own columns
        compact = part["norm"].numpy()
        for k, s in enumerate(mine):
            np.testing.assert_array_equal(compact[:, part["pred_idx_mine"][k]], norm[:, full["pred_idx"][s]])
            np.testing.assert_array_equal(compact[:, part["targ_idx_mine"][k]], norm[:, full["targ_idx"][s]])
        np.testing.assert_array_equal(part["train_rows"], full["train_rows"])
    assert seen == set(range(len(full["pred_idx"])))                              # every sub-network has an owner
    emu = bench.build_workload("tiny", "cpu", emulate=(1, world))
    assert emu["mine"] == bench.build_workload("tiny", "cpu", world, 1)["mine"]


def test_reference_arm_line_is_complete():
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--workload", "tiny", "--steps", "1",
                          "--warmup", "1", "--epochs", "2", "--ref-subnets", "2"], cwd=ROOT, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "cells*genes/s" and line["higher_is_better"] is True
    base = line["cpu_baseline"]
    assert base["kind"] == "port" and base["cores"] >= 1 and base["epoch_fraction"] == 1.0
    assert base["extrapolation_factor"] >= 1.0 and base["sampled_seconds"] > 0
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    # the value is N * G / (extrapolated fit + predict seconds)
    assert abs(line["value"] * line["ms_per_step"] * 1e-3 - 2000 * 1500) < 1.0
    # both arms print the same ``config`` object (what is computed); how an arm computes it goes under "arm"
    assert line["config"] == bench.workload_config(bench.build_workload("tiny", "cpu"), 2)
    assert set(line["config"]) == {"workload", "epochs_per_step", "batch_size", "sub_networks", "hidden", "sub_outputdim",
                                   "predictors_per_subnet", "adam_steps_per_epoch", "l2"}
    assert line["arm"]["extrapolation_factor"] == base["extrapolation_factor"]


def test_config_of_a_rank_is_the_config_of_the_job():
    """Sharded runs: the per-rank workload describes the whole job in ``config`` (rank 0 prints it)."""
    full = bench.workload_config(bench.build_workload("tiny", "cpu"), 20)
    part = bench.workload_config(bench.build_workload("tiny", "cpu", 3, 0), 20)
    assert full == part
