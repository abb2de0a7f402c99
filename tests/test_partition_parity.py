"""Host-side partitioning must reproduce the reference exactly (same np.random stream, same tie-breaking):
against golden vectors minted from the reference's own functions (scripts/make_golden.py) and, where
/root/reference is mounted, against those functions imported live (filter_genes multinet.py:312-331, setTargets
:333-342, setPredictors :344-365, get_distance_matrix :20-34)."""
import os
import sys
import types

import numpy as np
import pandas as pd
import pytest

from deepimpute_b200 import partition
from deepimpute_b200.multinet import MultiNet, get_distance_matrix
from conftest import REFERENCE, synthetic_counts


def host_partition(raw, seed, NN_lim=None, minVMR=0.5, ntop=5, sub_outputdim=512, n_pred=None):
    """The part of MultiNet.fit that runs before the engine is built (no GPU needed)."""
    net = MultiNet(seed=seed, sub_outputdim=sub_outputdim, ncores=1)
    np.random.seed(seed)
    ranked, metric = partition.rank_genes(raw)
    genes = partition.choose_genes(ranked, metric, sub_outputdim, minVMR, NN_lim)
    cand = partition.candidate_predictors(raw, n_pred)
    corr = partition.abs_correlation(raw.values, cand)
    net._set_partition(raw.columns, raw.values, genes, cand, corr, ntop, "random")
    np.random.seed(seed)
    train_rows, test_rows = partition.split_cells(raw.shape[0], labels=raw.index.values)
    cols = raw.columns
    return dict(targets=cols.get_indexer(net.targets.reshape(-1)).reshape(net.targets.shape),
                predictors=[cols.get_indexer(p) for p in net.predictors], test_rows=test_rows,
                train_rows=train_rows)


def assert_same(got, want):
    np.testing.assert_array_equal(got["targets"], want["targets"])
    assert [len(p) for p in got["predictors"]] == [len(p) for p in want["predictors"]]
    for a, b in zip(got["predictors"], want["predictors"]):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(got["test_rows"], want["test_rows"])
    np.testing.assert_array_equal(got["train_rows"], want["train_rows"])


def test_default_seed1234_matches_golden(test_counts, golden_partition):
    want = golden_partition("default_seed1234")
    got = host_partition(test_counts, 1234)
    assert [len(p) for p in got["predictors"]] == [639, 592, 592, 594, 555, 631]      # SURVEY.md 8c probe values
    assert got["targets"].shape == (6, 512) and len(np.unique(got["targets"])) == 3000
    assert len(got["test_rows"]) == 25 and len(got["train_rows"]) == 475
    assert_same(got, want)


def test_reference_multinet_test_setup_matches_golden(test_counts, golden_partition):
    # reference tests/multinet_test.py:14-29: the 1300 genes with the highest 99th percentile, seed 123
    sub = test_counts[test_counts.quantile(.99).sort_values(ascending=False).index[0:1300]]
    got = host_partition(sub, 123)
    assert [len(p) for p in got["predictors"]] == [303, 296, 284]
    assert_same(got, golden_partition("multinet_test"))


def test_reference_deepimpute_test_setup_matches_golden(test_counts, golden_partition):
    # reference tests/deepImpute_test.py:8-24: limit=1000, seed 1234
    got = host_partition(test_counts, 1234, NN_lim=1000)
    assert [len(p) for p in got["predictors"]] == [408, 405, 381]
    assert len(np.unique(got["targets"])) == 1342
    assert_same(got, golden_partition("deepimpute_test"))


def test_filter_genes_quirk_adds_a_whole_filler_network():
    # multinet.py:323-327: pad = O - (len % O) is O, not 0, when len is already a multiple of O
    raw = synthetic_counts(60, 40, seed=1)
    np.random.seed(0)
    ranked, metric = partition.rank_genes(raw)
    assert len(partition.choose_genes(ranked, metric, 8, 0.0, limit=16)) == 24
    assert len(partition.choose_genes(ranked, metric, 8, 0.0, limit=13)) == 24
    assert len(partition.choose_genes(ranked, metric, 8, 0.0, limit=17)) == 32


# ---- live comparison with the reference's own functions (build container only) --------------------------------
def _import_reference():
    if not os.path.isdir(REFERENCE):
        pytest.skip("/root/reference not mounted")
    for name in ["tensorflow", "tensorflow.keras", "keras", "keras.backend", "keras.models", "keras.layers",
                 "keras.callbacks", "keras.losses"]:
        m = types.ModuleType(name)
        m.__dict__.update(dict(backend=None, Model=None, model_from_json=None, Dense=None, Dropout=None,
                               Input=None, EarlyStopping=None))
        sys.modules.setdefault(name, m)
    sys.modules["tensorflow"].keras = sys.modules["tensorflow.keras"]
    sys.modules["keras"].losses = sys.modules["keras.losses"]
    sys.modules["keras"].backend = sys.modules["keras.backend"]
    if REFERENCE not in sys.path:
        sys.path.append(REFERENCE)
    import deepimpute.multinet as ref
    return ref


def _reference_partition(ref, raw, seed, sub_outputdim, NN_lim, minVMR, ntop):
    net = ref.MultiNet.__new__(ref.MultiNet)
    net.sub_outputdim, net.seed = sub_outputdim, seed
    np.random.seed(seed)
    gene_metric = (raw.var() / (1 + raw.mean())).sort_values(ascending=False)
    gene_metric = gene_metric[gene_metric > 0]
    genes = net.filter_genes(gene_metric, minVMR, NN_lim=NN_lim)
    cov = ref.get_distance_matrix(raw, n_pred=None)
    net.setTargets(raw.reindex(columns=genes), mode="random")
    net.setPredictors(cov, ntop=ntop)
    np.random.seed(seed)
    test_cells = np.random.choice(raw.index, int(0.05 * raw.shape[0]), replace=False)
    train_cells = np.setdiff1d(raw.index, test_cells)
    cols, idx = raw.columns, raw.index
    return dict(targets=cols.get_indexer(net.targets.reshape(-1)).reshape(net.targets.shape),
                predictors=[cols.get_indexer(p) for p in net.predictors],
                test_rows=idx.get_indexer(test_cells), train_rows=idx.get_indexer(train_cells)), cov


@pytest.mark.parametrize("seed,shape,O,NN_lim,ntop", [(1, (120, 90), 16, None, 5), (7, (64, 200), 32, 100, 3),
                                                       (42, (300, 70), 8, "auto", 5)])
def test_live_reference_partition_on_synthetic(seed, shape, O, NN_lim, ntop):
    ref = _import_reference()
    raw = synthetic_counts(*shape, seed=seed)
    want, cov = _reference_partition(ref, raw, seed, O, NN_lim, 0.5, ntop)
    got = host_partition(raw, seed, NN_lim=NN_lim, ntop=ntop, sub_outputdim=O)
    assert_same(got, want)
    mine = get_distance_matrix(raw)
    assert list(mine.columns) == list(cov.columns)
    np.testing.assert_allclose(mine.values, cov.values, rtol=0, atol=1e-12)


def test_live_reference_named_methods(test_counts):
    """MultiNet.filter_genes / setTargets / setPredictors keep the reference's signatures and results."""
    ref = _import_reference()
    raw = test_counts.iloc[:, :700]
    rnet = ref.MultiNet.__new__(ref.MultiNet)
    rnet.sub_outputdim, rnet.seed = 128, 5
    mine = MultiNet(sub_outputdim=128, seed=5, ncores=1)
    metric = (raw.var() / (1 + raw.mean())).sort_values(ascending=False)
    metric = metric[metric > 0]
    np.random.seed(5)
    g_ref = rnet.filter_genes(metric, 0.5, NN_lim=300)
    np.random.seed(5)
    g_mine = mine.filter_genes(metric, 0.5, NN_lim=300)
    assert list(g_ref) == list(g_mine)
    np.random.seed(6)
    rnet.setTargets(raw.reindex(columns=g_ref), mode="random")
    np.random.seed(6)
    mine.setTargets(raw.reindex(columns=g_mine), mode="random")
    np.testing.assert_array_equal(rnet.targets, mine.targets)
    cov = ref.get_distance_matrix(raw)
    rnet.setPredictors(cov, ntop=5)
    mine.setPredictors(cov, ntop=5)
    for a, b in zip(rnet.predictors, mine.predictors):
        assert list(a) == list(b)
    # mode='progressive' (multinet.py:337-338) is a plain reshape; the reference's own line fails on pandas 3
    # (Arrow-backed column labels have no reshape), so the expectation is written out
    mine.setTargets(raw.reindex(columns=g_mine), mode="progressive")
    np.testing.assert_array_equal(mine.targets, np.asarray(list(g_mine), dtype=object).reshape(-1, 128))


def test_statistics_based_filters_equal_the_pandas_ones(test_counts):
    """``rank_genes_from_stats`` / ``candidate_predictors_from_stats`` (fed by ``di_gene_stats`` on large inputs) apply
    the same rules as the pandas forms: with pandas' own mean / var they reproduce ranking and candidates exactly."""
    from deepimpute_b200 import partition
    raw = test_counts
    mean, var = raw.mean().values, raw.var().values
    ranked, metric = partition.rank_genes(raw)
    ranked2, metric2 = partition.rank_genes_from_stats(mean, var)
    np.testing.assert_array_equal(ranked, ranked2)
    np.testing.assert_array_equal(metric, metric2)
    for n_pred in (None, 700):
        np.testing.assert_array_equal(partition.candidate_predictors(raw, n_pred),
                                      partition.candidate_predictors_from_stats(mean, var, n_pred))
    # a gene that is zero everywhere has mean 0 and variance 0: std / mean = NaN -> not a candidate, metric 0 -> not ranked
    raw0 = raw.copy()
    raw0.iloc[:, 5] = 0.0
    m0, v0 = raw0.mean().values, raw0.var().values
    assert 5 not in partition.candidate_predictors_from_stats(m0, v0) and 5 not in partition.rank_genes_from_stats(m0, v0)[0]
    np.testing.assert_array_equal(partition.candidate_predictors(raw0), partition.candidate_predictors_from_stats(m0, v0))


# ---- n_pred: candidate predictors capped at the n_pred genes with the largest std/mean (multinet.py:25-33) -----------
def _reference_partition_n_pred(ref, raw, seed, sub_outputdim, NN_lim, ntop, n_pred):
    net = ref.MultiNet.__new__(ref.MultiNet)
    net.sub_outputdim, net.seed = sub_outputdim, seed
    np.random.seed(seed)
    gene_metric = (raw.var() / (1 + raw.mean())).sort_values(ascending=False)
    gene_metric = gene_metric[gene_metric > 0]
    genes = net.filter_genes(gene_metric, 0.5, NN_lim=NN_lim)
    cov = ref.get_distance_matrix(raw, n_pred=n_pred)
    net.setTargets(raw.reindex(columns=genes), mode="random")
    net.setPredictors(cov, ntop=ntop)                # KeyError when a target is not among the n_pred candidates
    cols = raw.columns
    return dict(targets=cols.get_indexer(net.targets.reshape(-1)).reshape(net.targets.shape),
                predictors=[cols.get_indexer(p) for p in net.predictors])


def test_live_reference_n_pred_where_the_reference_does_not_raise():
    """With every target inside the candidate set the reference's n_pred path works (multinet.py:356-358) and this
    port must pick the same predictors.  n_pred = number of genes with std/mean > 0 keeps every gene a candidate but
    takes the n_pred branch (candidates in std/mean order instead of column order)."""
    ref = _import_reference()
    raw = synthetic_counts(150, 80, seed=9)
    cv = raw.std() / raw.mean()
    n_pred = int((cv > 0).sum())
    want = _reference_partition_n_pred(ref, raw, 3, 16, None, 5, n_pred)
    got = host_partition(raw, 3, sub_outputdim=16, n_pred=n_pred)
    np.testing.assert_array_equal(got["targets"], want["targets"])
    for a, b in zip(got["predictors"], want["predictors"]):
        np.testing.assert_array_equal(a, b)


def test_n_pred_semantic_where_the_reference_raises():
    """Targets outside the n_pred candidates make the reference raise KeyError (cov.loc[targets] has no such rows,
    SURVEY.md 7.2).  Defined semantic here: rows = ALL targets, columns = the n_pred candidates;
    checked against a brute-force float64 restatement.  The whole matrix is never widened to float64."""
    raw = synthetic_counts(200, 120, seed=4).astype(np.float32)
    n_pred = 30
    got = host_partition(raw, 5, sub_outputdim=16, n_pred=n_pred)
    cand = partition.candidate_predictors(raw, n_pred)
    assert len(cand) == n_pred
    x = raw.values.astype(np.float64)
    c = np.abs(np.corrcoef(x.T))
    labels = np.asarray(raw.columns, dtype=object)
    for t, p in zip(got["targets"], got["predictors"]):
        assert not np.isin(t, cand).all()                              # the case the reference cannot handle
        keep = np.array(sorted(np.setdiff1d(cand, t), key=lambda g: labels[g]))
        top = np.argsort(-c[np.ix_(t, keep)], axis=1)[:, :5].ravel()
        np.testing.assert_array_equal(p, pd.unique(keep[top]))
        assert np.isin(p, cand).all() and len(p) <= n_pred


def test_user_gene_list_longer_than_one_subnetwork_is_padded_to_a_multiple():
    raw = synthetic_counts(60, 64, seed=1)
    ranked, _ = partition.rank_genes(raw)
    np.random.seed(0)
    user = np.arange(20)
    genes = partition.pad_user_genes(user, ranked, 16)
    assert len(genes) == 32 and list(genes[:20]) == list(user)         # every requested gene is kept
    assert len(partition.pad_user_genes(np.arange(10), ranked, 16)) == 16
    assert len(partition.pad_user_genes(np.arange(32), ranked, 16)) == 32
