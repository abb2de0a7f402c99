"""SURVEY.md 8f, the statistics behind both gene filters on the GPU (``di_gene_stats``): per-gene mean / variance against
pandas, the ranking and candidate rules built on them, and ``MultiNet(stats_engine="gpu")`` against the host route.
(Named to run after the other GPU suites: these are the newest tests of round 1.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gene_stats_match_pandas(test_counts, dtype):
    """di_gene_stats: float64 mean / variance (ddof 1) per gene against pandas, and the two gene filters built on them."""
    from deepimpute_b200 import partition
    raw = test_counts.astype(dtype)
    mean, var, ms = partition.gene_stats_gpu(raw.values)
    ref = test_counts.astype(np.float64)
    np.testing.assert_allclose(mean, ref.mean().values, rtol=1e-13, atol=0)
    np.testing.assert_allclose(var, ref.var().values, rtol=1e-12, atol=0)
    assert ms > 0
    ranked, metric = partition.rank_genes(ref)
    ranked_gpu, metric_gpu = partition.rank_genes_from_stats(mean, var)
    np.testing.assert_allclose(metric_gpu, metric, rtol=1e-12)          # same sorted metric values ...
    assert sorted(ranked_gpu) == sorted(ranked)                          # ... for the same genes
    moved = ranked_gpu != ranked                                         # order may differ only inside groups of ties
    if moved.any():
        by_gene = dict(zip(ranked, metric))
        np.testing.assert_allclose([by_gene[g] for g in ranked_gpu[moved]], metric[moved], rtol=1e-12)
    assert sorted(partition.candidate_predictors_from_stats(mean, var)) == sorted(partition.candidate_predictors(ref))
    with pytest.raises(RuntimeError):
        partition.gene_stats_gpu(np.zeros((1, 4), np.float32))          # a variance needs two cells


def test_multinet_with_gpu_statistics_keeps_the_partition(test_counts):
    from deepimpute_b200.multinet import MultiNet
    a = MultiNet(seed=1234, ncores=1, max_epochs=1, verbose=0, stats_engine="gpu").fit(test_counts)
    b = MultiNet(seed=1234, ncores=1, max_epochs=1, verbose=0, stats_engine="host").fit(test_counts)
    assert a.timings["stats_engine"] == "gpu" and b.timings["stats_engine"] == "host"
    assert a.targets.shape == b.targets.shape and len(a.predictors) == len(b.predictors)
    if np.array_equal(a._ranked, b._ranked):                            # no tie was ordered differently: identical partition
        np.testing.assert_array_equal(a.targets, b.targets)
        assert all(list(p) == list(q) for p, q in zip(a.predictors, b.predictors))
    else:
        common = len(set(a.targets.flatten()) & set(b.targets.flatten()))
        assert common >= 0.99 * len(set(b.targets.flatten()))
