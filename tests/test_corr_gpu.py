"""SURVEY.md 8f row 1 on the GPU: ``di_corr_topk`` against the host path, which is pinned to the reference's own
``get_distance_matrix`` / ``setPredictors`` (tests/test_partition_parity.py).  The device works in fp32, the
reference in float64: values agree to 2e-5 and the selected predictors are identical except where two candidates
of a target are tied to within that error (bounded below)."""
import numpy as np
import pytest

from conftest import synthetic_counts
from deepimpute_b200 import MultiNet, _lib, partition

pytestmark = pytest.mark.gpu


def host_and_gpu(raw, seed, O, NN_lim=None, ntop=5):
    np.random.seed(seed)
    ranked, metric = partition.rank_genes(raw)
    genes = partition.choose_genes(ranked, metric, O, 0.5, NN_lim)
    targets = partition.assign_targets(genes, O)
    cand = partition.candidate_predictors(raw)
    labels = np.asarray(raw.columns, dtype=object)[cand]
    corr = partition.abs_correlation(raw.values, cand)
    where = np.full(raw.shape[1], -1)
    where[cand] = np.arange(len(cand))
    host = partition.choose_predictors(targets, cand, labels, lambda i, t: corr[where[t]], ntop)
    gpu, ms = partition.choose_predictors_gpu(raw.values, targets, cand, labels, ntop)
    return targets, cand, labels, corr, where, host, gpu, ms


def check_same_up_to_ties(targets, cand, labels, corr, where, host, gpu, ntop, tol=5e-6):
    """Every sub-network: identical ordered predictor lists, or every gene picked by only one side is tied (within
    tol) with the weakest pick of some target of that sub-network."""
    n_diff = 0
    for s, (h, g) in enumerate(zip(host, gpu)):
        if len(h) == len(g) and (h == g).all():
            continue
        n_diff += 1
        only = np.setxor1d(h, g)
        rows = corr[where[targets[s]]]                                  # [O, n_cand] |r| of this sub-network's targets
        keep = ~np.isin(cand, targets[s])
        kth = -np.sort(-rows[:, keep], axis=1)[:, ntop - 1]             # weakest selected correlation per target
        for gene in only:
            gap = np.abs(rows[:, where[gene]] - kth).min()
            assert gap < tol, "gene {} of sub-network {} differs without a tie (gap {:.2e})".format(gene, s, gap)
    return n_diff


def test_values_match_numpy_corrcoef():
    import ctypes as C
    raw = synthetic_counts(700, 300, seed=2)
    cand = np.arange(300, dtype=np.int32)
    targ = np.arange(256, dtype=np.int32).reshape(2, 128)
    top = np.empty((2, 128, 5), np.int32)
    val = np.empty((2, 128, 5), np.float32)
    raw32 = np.ascontiguousarray(raw.values, dtype=np.float32)
    lib = _lib.load()
    rc = lib.di_corr_topk(0, _lib.f32(raw32), 700, 300, _lib.i32(cand), 300, _lib.i32(targ), 2, 128, 5, _lib.i32(top),
                          _lib.f32(val), None)
    assert rc == 0, lib.di_corr_last_error()
    want = np.abs(np.nan_to_num(np.corrcoef(raw.values.T)))
    for s in range(2):
        for o in range(128):
            t = targ[s, o]
            np.testing.assert_allclose(val[s, o], want[t, top[s, o]], atol=2e-5)
            assert (np.diff(val[s, o]) <= 0).all()                       # descending
            assert not np.isin(top[s, o], targ[s]).any()                 # own targets excluded
            masked = want[t].copy()
            masked[targ[s]] = -1
            assert val[s, o, -1] >= np.sort(masked)[-5] - 2e-5           # nothing better was missed
    # bad arguments are rejected
    assert lib.di_corr_topk(0, _lib.f32(raw32), 700, 300, _lib.i32(cand), 300, _lib.i32(targ), 2, 128, 9, _lib.i32(top),
                            None, None) == 1


def test_selection_matches_host_on_the_example_matrix(test_counts, golden_partition):
    out = host_and_gpu(test_counts, 1234, 512)
    targets, cand, labels, corr, where, host, gpu, ms = out
    want = golden_partition("default_seed1234")["predictors"]
    for a, b in zip(host, want):
        np.testing.assert_array_equal(a, b)                              # host path == reference (golden)
    n_diff = check_same_up_to_ties(targets, cand, labels, corr, where, host, gpu, 5)
    print("test.csv: {} of {} sub-networks differ at ties; device {:.1f} ms".format(n_diff, len(host), ms))
    same = sum(len(np.intersect1d(h, g)) for h, g in zip(host, gpu)) / sum(len(h) for h in host)
    assert same > 0.995


@pytest.mark.parametrize("shape,O,seed", [((900, 700), 64, 3), ((333, 1030), 128, 8)])
def test_selection_matches_host_on_synthetic(shape, O, seed):
    raw = synthetic_counts(*shape, seed=seed)
    targets, cand, labels, corr, where, host, gpu, ms = host_and_gpu(raw, seed, O, NN_lim=400)
    check_same_up_to_ties(targets, cand, labels, corr, where, host, gpu, 5)


def test_multinet_with_gpu_predictor_engine(test_counts):
    net = MultiNet(seed=1234, ncores=1, max_epochs=2, verbose=0, predictor_engine="gpu")
    net.fit(test_counts)
    assert net.timings["predictor_engine"] == "gpu" and net.timings["predictor_selection_device_ms"] > 0
    sizes = [len(p) for p in net.predictors]
    assert np.abs(np.array(sizes) - np.array([639, 592, 592, 594, 555, 631])).max() <= 2
    assert net.predict(test_counts).shape == test_counts.shape
