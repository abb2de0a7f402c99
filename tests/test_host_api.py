"""Host-side mirror of the reference interface: signatures, flag defaults, and the pandas post-processing of
``predict`` (multinet.py:282-310), run on CPU with the oracle-backed stand-in engine (tests/fake_engine.py)."""
import inspect

import numpy as np
import pandas as pd
import pytest

from conftest import synthetic_counts
from deepimpute_b200 import MultiNet, deepImpute, inspect_data, wMSE
from deepimpute_b200.parser import parse_args
from fake_engine import FakeEngine


class CpuMultiNet(MultiNet):
    def _make_engine(self, inputdims, **kw):
        return FakeEngine(inputdims, **kw)


def test_constructor_and_method_signatures_match_the_reference():
    # reference multinet.py:67-79, :169-178, :266-269
    p = inspect.signature(MultiNet.__init__).parameters
    ref = [("learning_rate", 1e-4), ("batch_size", 64), ("max_epochs", 500), ("patience", 5), ("ncores", -1),
           ("loss", "wMSE"), ("output_prefix", None), ("sub_outputdim", 512), ("verbose", 1), ("seed", 1234),
           ("architecture", None)]
    assert list(p)[1:12] == [k for k, _ in ref]
    for k, v in ref:
        if k != "output_prefix":
            assert p[k].default == v
    f = inspect.signature(MultiNet.fit).parameters
    assert list(f)[1:] == ["raw", "cell_subset", "NN_lim", "genes_to_impute", "n_pred", "ntop", "minVMR", "mode"]
    assert (f["cell_subset"].default, f["ntop"].default, f["minVMR"].default, f["mode"].default) == (1, 5, .5, "random")
    q = inspect.signature(MultiNet.predict).parameters
    assert list(q)[1:] == ["raw", "imputed_only", "policy"] and q["policy"].default == "restore"
    net = MultiNet(ncores=2)
    net.loadDefaultArchitecture()
    assert net.NN_parameters["architecture"] == [{"type": "dense", "neurons": 256, "activation": "relu"},
                                                 {"type": "dropout", "rate": 0.2}]


def test_cli_defaults_match_reference_parser():
    a = parse_args(["in.csv"])                                     # reference parser.py:3-95
    assert (a.output, a.cores, a.cell_axis, a.limit, a.minVMR, a.subset) == ("./imputed.csv", -1, "rows", "auto", .5, 1)
    assert (a.learning_rate, a.batch_size, a.max_epochs, a.hidden_neurons) == (0.0005, 64, 300, 300)
    assert (a.dropout_rate, a.output_neurons, a.n_pred, a.policy) == (0.2, 512, None, "restore")
    b = parse_args(["in.csv", "-o", "x.csv", "--cell-axis", "columns", "--limit", "2000", "--n_pred", "50"])
    assert (b.output, b.cell_axis, b.limit, b.n_pred) == ("x.csv", "columns", "2000", 50)


def test_inspect_data_rejects_bad_input():
    ok = synthetic_counts(10, 6)
    inspect_data(ok)
    with pytest.raises(ValueError, match="duplicated cell"):
        inspect_data(pd.concat([ok, ok.iloc[:1]]))
    dup = ok.copy()
    dup.columns = ["g0"] * 2 + list(ok.columns[2:])
    with pytest.raises(ValueError, match="duplicated gene"):
        inspect_data(dup)
    with pytest.raises(ValueError, match="log-transformed"):
        inspect_data(np.log1p(ok).clip(upper=5))


def test_wmse_numpy_form():
    y = np.array([[0.0, 2.0], [1.0, 0.0]])
    p = np.array([[5.0, 1.0], [0.0, 7.0]])
    assert wMSE(y, p) == pytest.approx((2.0 * 1.0 + 1.0 * 1.0) / 4)
    assert wMSE(y, p, binary=True) == pytest.approx((1.0 + 1.0) / 4)


def test_unsupported_topologies_fail_loudly():
    raw = synthetic_counts(40, 30)
    with pytest.raises(NotImplementedError):
        CpuMultiNet(ncores=1, architecture=[{"type": "dense", "neurons": 8, "activation": "relu"},
                                            {"type": "dense", "neurons": 8, "activation": "relu"}]).build([4])
    with pytest.raises(NotImplementedError):
        CpuMultiNet(ncores=1, loss="mean_squared_error").build([4])
    del raw


@pytest.fixture(scope="module")
def fitted():
    raw = synthetic_counts(120, 90, seed=3)
    net = CpuMultiNet(ncores=1, sub_outputdim=16, max_epochs=3, seed=11, verbose=0,
                      architecture=[{"type": "dense", "neurons": 12, "activation": "relu"},
                                    {"type": "dropout", "rate": 0.2}])
    net.fit(raw, NN_lim=40, minVMR=0.0)
    return raw, net


def test_fit_sets_reference_attributes(fitted):
    raw, net = fitted
    # ceil(40/16) = 3 nets = 48 genes, then the filler quirk adds 16 - 48 % 16 = 16 more -> 4 nets (multinet.py:323)
    assert net.targets.shape == (4, 16)
    assert len(net.predictors) == net.targets.shape[0]
    assert net.trained_epochs == 3
    assert set(net.test_metrics) == {"correlation", "MSE"}
    assert len(net.test_cells) == 6 and len(net.train_cells) == 114
    assert list(net.train_cells) == sorted(net.train_cells)          # np.setdiff1d order (multinet.py:229)


def test_predict_postprocessing_matches_a_pandas_restatement(fitted):
    raw, net = fitted
    out = net.predict(raw, policy="restore")
    assert out.shape == raw.shape and list(out.columns) == list(raw.columns) and list(out.index) == list(raw.index)
    assert out.values.dtype == np.float64
    # restatement of multinet.py:271-305 with pandas (groupby(axis=1) spelled as a transpose for pandas 3)
    norm = np.log1p(raw)
    cols = raw.columns
    pred_idx = [cols.get_indexer(p) for p in net.predictors]
    targ = net.targets.flatten()
    net.engine.set_data(norm.values.astype(np.float32), pred_idx, cols.get_indexer(targ).reshape(net.targets.shape))
    predicted = pd.DataFrame(net.engine.predict(), index=raw.index, columns=targ)
    predicted = predicted.T.groupby(level=0).mean().T
    not_predicted = norm.drop(targ, axis=1)
    imputed = pd.concat([predicted, not_predicted], axis=1).loc[raw.index, raw.columns].values.astype(np.float64)
    imputed[(imputed > 2 * norm.values.max()) | np.isnan(imputed)] = 0
    imputed = np.expm1(imputed)
    mask = raw.values > 0
    restored = imputed.copy()
    restored[mask] = raw.values[mask]
    np.testing.assert_allclose(out.values, restored, rtol=1e-6)
    # policy 'max' and imputed_only
    out_max = net.predict(raw, policy="max")
    np.testing.assert_allclose(out_max.values, np.maximum(imputed, raw.values), rtol=1e-6)
    only = net.predict(raw, imputed_only=True)
    assert sorted(only.columns) == sorted(set(targ))
    # every originally positive count survives 'restore'
    assert np.array_equal(out.values[mask], raw.values[mask])


def test_deepimpute_entry_point_runs_on_csv(tmp_path, monkeypatch):
    import sys
    entry = sys.modules["deepimpute_b200.deepImpute"]     # the package attribute of that name is the function
    monkeypatch.setattr(entry, "MultiNet", CpuMultiNet)
    raw = synthetic_counts(80, 60, seed=5)
    path = tmp_path / "counts.csv"
    raw.to_csv(path)
    out = deepImpute(inputFile=str(path), output=None, max_epochs=2, hidden_neurons=8, output_neurons=16,
                     limit="20", cores=1)
    assert out.shape == raw.shape
    raw.T.to_csv(path)
    out2 = entry.deepImpute(inputFile=str(path), output=str(tmp_path / "o.csv"), max_epochs=1, hidden_neurons=8,
                            output_neurons=16, limit="20", cores=1, cell_axis="columns")
    assert out2 is None and (tmp_path / "o.csv").exists()


def test_unknown_genes_to_impute_raise_like_the_reference():
    # the reference raises KeyError (reindex / .loc, multinet.py:213, :356); get_indexer's -1 must not become "last column"
    raw = synthetic_counts(60, 40, seed=2)
    net = CpuMultiNet(ncores=1, sub_outputdim=8, max_epochs=1, verbose=0,
                      architecture=[{"type": "dense", "neurons": 4, "activation": "relu"}])
    with pytest.raises(KeyError, match="NOPE"):
        net.fit(raw, genes_to_impute=["g1", "g2", "NOPE"], minVMR=0.0)


def test_sharded_fit_needs_a_seed():
    from deepimpute_b200.parallel import ShardContext
    raw = synthetic_counts(60, 40, seed=2)
    net = CpuMultiNet(ncores=1, sub_outputdim=8, max_epochs=1, verbose=0, seed=None, shard=ShardContext(0, 2))
    with pytest.raises(ValueError, match="seed"):
        net.fit(raw)


@pytest.mark.parametrize("cell_subset, n_cells", [(0.5, 60), (80, 80)])
def test_cell_subset_draws_the_cells_the_reference_draws(cell_subset, n_cells):
    """``fit(cell_subset=...)`` (multinet.py:185-189): a fraction or a count of cells, drawn by ``DataFrame.sample`` from
    the numpy stream seeded just before -- so the same seed picks the same cells, and everything downstream (split,
    training) sees only them."""
    raw = synthetic_counts(120, 60, seed=8)
    np.random.seed(5)
    want = raw.sample(frac=cell_subset) if cell_subset < 1 else raw.sample(cell_subset)
    net = CpuMultiNet(ncores=1, sub_outputdim=16, max_epochs=1, seed=5, verbose=0,
                      architecture=[{"type": "dense", "neurons": 8, "activation": "relu"}])
    net.fit(raw, cell_subset=cell_subset, NN_lim=30, minVMR=0.0)
    used = np.concatenate([net.train_cells, net.test_cells])
    assert len(used) == n_cells and set(used) == set(want.index)
    assert len(net.test_cells) == int(0.05 * n_cells)
    out = net.predict(raw)                                   # predict is free to see every cell again
    assert out.shape == raw.shape and np.isfinite(out.values).all()
