"""Mint the golden vectors under tests/golden/ from the REFERENCE's own host functions.

Run in the build container only (needs /root/reference):  python scripts/make_golden.py
TensorFlow/Keras are not installed, so they are stubbed in sys.modules; only the reference's pure
numpy/pandas functions are executed (multinet.py:20-34 get_distance_matrix, :312-331 filter_genes,
:333-342 setTargets, :344-365 setPredictors) in the exact order MultiNet.fit calls them (:180-229).
Outputs:
  tests/golden/test_counts.npz      the example matrix (examples/test.csv) as int32 + labels, so GPU-box tests
                                    can run config 1 without /root/reference
  tests/golden/partition_*.npz      targets / predictors / test cells for three set-ups
"""
import os
import sys
import types

import numpy as np
import pandas as pd

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    for name in ["tensorflow", "tensorflow.keras", "keras", "keras.backend", "keras.models", "keras.layers",
                 "keras.callbacks", "keras.losses"]:
        m = types.ModuleType(name)
        m.__dict__.update(dict(backend=None, Model=None, model_from_json=None, Dense=None, Dropout=None,
                               Input=None, EarlyStopping=None))
        sys.modules[name] = m
    sys.modules["tensorflow"].keras = sys.modules["tensorflow.keras"]
    sys.modules["keras"].losses = sys.modules["keras.losses"]
    sys.modules["keras"].backend = sys.modules["keras.backend"]
    sys.path.insert(0, REF)
    import deepimpute.multinet as ref
    return ref


def reference_partition(ref, raw, seed, sub_outputdim=512, NN_lim=None, minVMR=0.5, ntop=5, n_pred=None):
    """The host-side part of reference MultiNet.fit (multinet.py:180-229), nothing else."""
    net = ref.MultiNet.__new__(ref.MultiNet)
    net.sub_outputdim, net.seed = sub_outputdim, seed
    np.random.seed(seed)
    gene_metric = (raw.var() / (1 + raw.mean())).sort_values(ascending=False)
    gene_metric = gene_metric[gene_metric > 0]
    genes = net.filter_genes(gene_metric, minVMR, NN_lim=NN_lim)
    cov = ref.get_distance_matrix(raw, n_pred=n_pred)
    net.setTargets(raw.reindex(columns=genes), mode="random")
    net.setPredictors(cov, ntop=ntop)
    np.random.seed(seed)
    norm_index = raw.index
    test_cells = np.random.choice(norm_index, int(0.05 * raw.shape[0]), replace=False)
    train_cells = np.setdiff1d(norm_index, test_cells)
    return net.targets, net.predictors, test_cells, train_cells


def save_case(name, raw, targets, predictors, test_cells, train_cells):
    cols, idx = raw.columns, raw.index
    np.savez_compressed(
        os.path.join(OUT, "partition_{}.npz".format(name)),
        targets=cols.get_indexer(targets.reshape(-1)).reshape(targets.shape).astype(np.int32),
        pred_flat=np.concatenate([cols.get_indexer(p) for p in predictors]).astype(np.int32),
        pred_len=np.asarray([len(p) for p in predictors], dtype=np.int32),
        test_rows=idx.get_indexer(test_cells).astype(np.int32),
        train_rows=idx.get_indexer(train_cells).astype(np.int32))
    print(name, "S =", len(predictors), "P_s =", [len(p) for p in predictors])


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = import_reference()
    raw = pd.read_csv(os.path.join(REF, "examples", "test.csv"), index_col=0)
    assert np.all(raw.values == np.round(raw.values))
    np.savez_compressed(os.path.join(OUT, "test_counts.npz"), counts=raw.values.astype(np.int32),
                        cells=np.asarray(list(raw.index), dtype="U"), genes=np.asarray(list(raw.columns), dtype="U"))
    # 1. defaults, seed 1234 (BASELINE.json configs[0])
    save_case("default_seed1234", raw, *reference_partition(ref, raw, 1234))
    # 2. reference tests/multinet_test.py:14-29: top-1300 genes by 99th percentile, seed 123
    sub = raw[raw.quantile(.99).sort_values(ascending=False).index[0:1300]]
    save_case("multinet_test", sub, *reference_partition(ref, sub, 123))
    # 3. reference tests/deepImpute_test.py:8-24: limit=1000, seed 1234
    save_case("deepimpute_test", raw, *reference_partition(ref, raw, 1234, NN_lim=1000))


if __name__ == "__main__":
    main()
