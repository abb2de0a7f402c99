"""End-to-end through the reference-facing API on the GPU: MultiNet.fit / predict and deepImpute() on the reference's
own example matrix (examples/test.csv, shipped as tests/golden/test_counts.npz), mirroring the reference's smoke
tests (tests/multinet_test.py:12-33, tests/deepImpute_test.py:8-32) -- plus what they do not check: parity of the
imputed matrix with the CPU oracle for the same seeds at a fixed epoch count."""
import numpy as np
import pytest

from conftest import synthetic_counts
from deepimpute_b200 import MultiNet, deepImpute
from deepimpute_b200.engine import epoch_permutation
from oracle.multinet_oracle import OracleNet, stage

pytestmark = pytest.mark.gpu


def test_reference_multinet_test_flow(test_counts):
    # reference tests/multinet_test.py:14-31
    raw = test_counts[test_counts.quantile(.99).sort_values(ascending=False).index[0:1300]]
    model = MultiNet(architecture=[{"type": "dense", "neurons": 150, "activation": "relu"},
                                   {"type": "dropout", "rate": 0.2}],
                     loss="wMSE", sub_outputdim=512, seed=123, ncores=2, max_epochs=30, verbose=0)
    model.fit(raw)
    out = model.predict(raw, policy="restore")
    assert out.shape == raw.shape
    assert [len(p) for p in model.predictors] == [303, 296, 284]
    assert np.isfinite(out.values).all() and (out.values >= 0).all()
    mask = raw.values > 0
    assert np.array_equal(out.values[mask], raw.values[mask])
    assert (out.values[~mask] > 0).mean() > 0.5                      # zeros of imputed genes get filled in
    assert model.test_metrics["correlation"] > 0.7
    assert 1 <= model.trained_epochs <= 30
    # a fresh object pointed at the same output_prefix predicts from the saved model (multinet.py:117-124)
    again = MultiNet(output_prefix=model.outputdir, sub_outputdim=512, ncores=2, verbose=0)
    out2 = again.predict(raw)
    np.testing.assert_allclose(out2.values, out.values, rtol=1e-6)


@pytest.mark.parametrize("math_mode", ["fp32", "tf32x3"])
def test_fit_predict_matches_oracle_at_fixed_epochs(test_counts, math_mode):
    """Config c1 (test.csv, defaults, seed 1234) for 5 epochs: imputed values within 1e-3 relative of the oracle --
    for the strict fp32 kernels and for the default tensor-core mode directly (not via the fp32 path) -- and the
    held-out metrics of multinet.py:252-262 equal to the oracle's own."""
    raw = test_counts
    net = MultiNet(seed=1234, ncores=1, max_epochs=5, patience=100, verbose=0, math_mode=math_mode)
    net.fit(raw)
    assert [len(p) for p in net.predictors] == [639, 592, 592, 594, 555, 631]
    got = net.predict(raw, policy="restore")

    cols = raw.columns
    norm = np.log1p(raw.values).astype(np.float32)
    pred_idx = [cols.get_indexer(p) for p in net.predictors]
    targ_idx = cols.get_indexer(net.targets.reshape(-1)).reshape(net.targets.shape)
    train_rows = raw.index.get_indexer(net.train_cells)
    test_rows = raw.index.get_indexer(net.test_cells)
    ref = OracleNet([len(p) for p in pred_idx], 256, 512, seed=1234)
    Xtr, Ytr = stage(norm, pred_idx, targ_idx, train_rows)
    Xte, Yte = stage(norm, pred_idx, targ_idx, test_rows)
    step, losses, vals = 0, [], []
    for e in range(5):
        loss, step = ref.train_epoch(Xtr, Ytr, epoch_permutation(1234, e, len(train_rows)), step)
        losses.append(loss)
        vals.append(ref.loss(Xte, Yte))
    np.testing.assert_allclose(net.history["loss"], losses, rtol=1e-4)
    np.testing.assert_allclose(net.history["val_loss"], vals, rtol=1e-4)
    want = np.hstack(ref.forward(stage(norm, pred_idx, targ_idx, np.arange(len(raw)))[0]))
    flat = targ_idx.reshape(-1)
    uniq, first = np.unique(flat, return_index=True)
    dup = np.setdiff1d(np.arange(len(flat)), first)
    single = np.setdiff1d(uniq, flat[dup])                 # genes predicted by exactly one output unit
    where = {g: i for i, g in enumerate(flat)}
    sel = np.array([where[g] for g in single])
    zero = raw.values[:, single] == 0
    imputed_ref = np.expm1(want[:, sel].astype(np.float64))
    rel = np.abs(got.values[:, single] - imputed_ref) / (np.abs(imputed_ref) + 1e-3)
    assert rel[zero].max() < 1e-3
    # held-out metrics (multinet.py:252-262): Pearson r and MSE over the originally non-zero entries of the test cells
    from scipy.stats import pearsonr
    y_true = np.hstack(Yte).reshape(-1)
    y_hat = np.hstack(ref.forward(Xte)).reshape(-1)
    seen = y_true > 0
    assert net.test_metrics["correlation"] == pytest.approx(pearsonr(y_true[seen], y_hat[seen])[0], rel=1e-5)
    assert net.test_metrics["MSE"] == pytest.approx(np.sum((y_true[seen] - y_hat[seen]) ** 2) / seen.sum(), rel=1e-4)


def test_tensor_core_modes_track_the_fp32_path(test_counts):
    """The same fit on the tensor-core paths.  Measured on a B200 (5 epochs, test.csv, imputed zeros, relative to the
    fp32 path which itself sits within 1e-3 of the oracle, see above):
      tf32   (operands truncated by the tensor core)          median 2.6e-3, 99th pct 4.5e-2, max 8.8e-2
      compensated forward GEMMs only                          median 1.1e-4, 99th pct 7.9e-3, max 2.9e-2
      compensated forward + dz1 GEMMs                         median 2.0e-5, 99th pct 6.3e-4, max 9.0e-4
      tf32x3 (all five GEMMs compensated, the default)        bounds asserted below
    so single-pass TF32 does NOT meet the 1e-3 target of the north star and is offered as an opt-in only."""
    from deepimpute_b200 import _lib
    if not _lib.load().di_math_mode_available(_lib.DI_MATH["tf32x3"]):
        pytest.skip("tensor-core kernels not built")
    raw = test_counts
    nets = {}
    for mode in ("fp32", "tf32", "tf32x3"):
        net = MultiNet(seed=1234, ncores=1, max_epochs=5, patience=100, verbose=0, math_mode=mode)
        net.fit(raw)
        nets[mode] = (net, net.predict(raw, policy="restore").values)
    zero = raw.values == 0
    stats = {}
    for mode in ("tf32", "tf32x3"):
        a, b = nets[mode][1][zero], nets["fp32"][1][zero]
        rel = np.abs(a - b) / (np.abs(b) + 1e-3)
        stats[mode] = (np.median(rel), np.quantile(rel, 0.99), rel.max())
        print("{} vs fp32 imputed values after 5 epochs: median rel {:.2e}, 99th pct {:.2e}, max {:.2e}".format(
            mode, *stats[mode]))
    np.testing.assert_allclose(nets["tf32"][0].history["loss"], nets["fp32"][0].history["loss"], rtol=1e-2)
    np.testing.assert_allclose(nets["tf32x3"][0].history["loss"], nets["fp32"][0].history["loss"], rtol=1e-3)
    np.testing.assert_allclose(nets["tf32x3"][0].history["val_loss"], nets["fp32"][0].history["val_loss"], rtol=1e-3)
    assert stats["tf32"][1] < 0.1
    assert stats["tf32x3"][2] < 1e-3 and stats["tf32x3"][1] < 2e-4 and stats["tf32x3"][0] < 5e-5
    assert abs(nets["tf32x3"][0].test_metrics["correlation"] - nets["fp32"][0].test_metrics["correlation"]) < 1e-4


def test_deepimpute_entry_point(tmp_path, test_counts):
    # reference tests/deepImpute_test.py:8-32 (limit 1000, hidden 300, lr 1e-4), fewer epochs
    path = tmp_path / "test.csv"
    test_counts.to_csv(path)
    out = deepImpute(inputFile=str(path), output=None, cores=1, cell_axis="rows", limit="1000", minVMR=0.5, subset=1,
                     learning_rate=1e-4, batch_size=64, max_epochs=3, hidden_neurons=300, dropout_rate=0.2,
                     output_neurons=512, n_pred=None, policy="restore")
    assert out.shape == test_counts.shape and np.isfinite(out.values).all()


def test_small_inputs_and_user_gene_list():
    raw = synthetic_counts(90, 70, seed=4)
    net = MultiNet(ncores=1, sub_outputdim=16, max_epochs=2, verbose=0,
                   architecture=[{"type": "dense", "neurons": 12, "activation": "relu"}, {"type": "dropout", "rate": 0.5}])
    net.fit(raw, genes_to_impute=list(raw.columns[:10]), minVMR=0.0)          # padded up to one sub-network
    assert net.targets.shape == (1, 16)
    out = net.predict(raw, imputed_only=True)
    assert out.shape[0] == 90 and set(raw.columns[:10]) <= set(out.columns)


def test_n_pred_route_on_the_gpu_matches_the_host_route(test_counts):
    """``fit(n_pred=...)`` caps the predictor candidates at the n_pred genes with the largest std/mean (multinet.py:25-29).
    The device route (di_corr_topk with that candidate list; rows = all targets) must choose the predictors the host
    route chooses, and the model must train and predict with them."""
    raw = test_counts.iloc[:, :1500]
    kw = dict(seed=7, ncores=1, max_epochs=2, patience=100, verbose=0, sub_outputdim=256)
    host = MultiNet(predictor_engine="host", **kw).fit(raw, n_pred=300, NN_lim=512)
    dev = MultiNet(predictor_engine="gpu", **kw).fit(raw, n_pred=300, NN_lim=512)
    assert dev.timings["predictor_engine"] == "gpu" and host.timings["predictor_engine"] == "host"
    np.testing.assert_array_equal(host.targets, dev.targets)
    same = sum(len(np.intersect1d(a, b)) for a, b in zip(host.predictors, dev.predictors))
    total = sum(len(p) for p in host.predictors)
    assert same >= 0.995 * total                                     # fp32 vs float64 correlations: ties only
    assert max(len(p) for p in dev.predictors) <= 300
    out = dev.predict(raw)
    assert out.shape == raw.shape and np.isfinite(out.values).all()


def test_integer_gene_labels_survive_save_and_load(tmp_path):
    """A frame whose gene labels are integers: a fresh object pointed at the saved model must find its genes again
    (the labels are stored with their dtype, not as strings)."""
    raw = synthetic_counts(120, 90, seed=6)
    raw.columns = np.arange(1000, 1000 + raw.shape[1])
    kw = dict(ncores=1, sub_outputdim=16, max_epochs=2, verbose=0, output_prefix=str(tmp_path),
              architecture=[{"type": "dense", "neurons": 12, "activation": "relu"}, {"type": "dropout", "rate": 0.2}])
    net = MultiNet(**kw)
    net.fit(raw, NN_lim=30, minVMR=0.0)
    out = net.predict(raw)
    again = MultiNet(**kw)
    out2 = again.predict(raw)
    np.testing.assert_allclose(out2.values, out.values, rtol=1e-6)
