"""Host orchestration layer: the reference's ``MultiNet`` with the Keras model swapped for the B200 engine.

Mirrors ``/root/reference/deepimpute/multinet.py`` (signatures ``:67-79``, ``:169-178``, ``:266-269``):
gene filtering, output-gene partitioning into ``sub_outputdim``-wide sub-networks, predictor selection by
correlation, the 5 % held-out split and the pandas post-processing all stay in Python and consume the
legacy ``np.random`` stream in exactly the reference's order (``:183``, ``:326``, ``:340``, ``:219``, ``:228``)
so that ``targets`` / ``predictors`` / the split are identical for the same seed.  Everything the reference
hands to Keras (``:226-253``, ``:276-280``) goes to :class:`deepimpute_b200.engine.Engine` instead, which runs
hand-written sm_100a kernels through the C-ABI in ``include/deepimpute_b200.h``.

Differences kept deliberately small and listed in DESIGN.md: bad input raises ``ValueError`` instead of
``exit(1)`` (``:46-58``); the duplicate-column mean of ``:284`` is written without ``groupby(axis=1)``
(removed in pandas 3); ``predict`` uses the weights resident in the engine (and falls back to the saved
``.npz`` when called on a fresh object) instead of re-loading Keras JSON/HDF5 (``:117-124``).
"""
import os
import tempfile
import warnings

import numpy as np
import pandas as pd
from scipy.stats import pearsonr

from . import partition


def _labels(index):
    """Index labels as a plain object ndarray (pandas 3 string indexes are Arrow-backed and lack ndarray indexing)."""
    return np.asarray(index, dtype=object)


def get_distance_matrix(raw, n_pred=None):
    """|Pearson r| between candidate predictor genes on RAW counts, as a labelled frame (``multinet.py:20-34``)."""
    cand = partition.candidate_predictors(raw, n_pred)
    labels = raw.columns[cand]
    return pd.DataFrame(partition.abs_correlation(raw.values, cand), index=labels, columns=labels)


def wMSE(y_true, y_pred, binary=False):
    """Weighted MSE of the reference (``multinet.py:36-41``), numpy form: mean over all axes of w*(y-yhat)^2."""
    y_true = np.asarray(y_true)
    y_pred = np.asarray(y_pred)
    weights = (y_true > 0).astype(np.float32) if binary else y_true
    return np.mean(weights * np.square(y_true - y_pred))


def inspect_data(data):
    """Input sanity checks of ``multinet.py:43-63``; raises ``ValueError`` where the reference calls ``exit(1)``."""
    if sum(data.index.duplicated()):
        raise ValueError("ERROR: duplicated cell labels. Please provide unique cell labels.")

    if sum(data.columns.duplicated()):
        raise ValueError("ERROR: duplicated gene labels. Please provide unique gene labels.")

    max_value = np.max(data.values)
    if max_value < 10:
        raise ValueError("ERROR: max value = {}. Is your data log-transformed? Please provide raw counts"
                         .format(max_value))

    print("Input dataset is {} cells (rows) and {} genes (columns)".format(*data.shape))
    print("First 3 rows and columns:")
    print(data.iloc[:3, :3])


class MultiNet:
    """Drop-in for ``deepimpute.multinet.MultiNet`` (reference ``multinet.py:65-375``).

    Extra keyword-only arguments (not in the reference): ``math_mode`` ("tf32x3" tensor-core kernels, default; "tf32"; or
    "fp32" CUDA-core kernels), ``postprocess`` ("gpu": fused device-side ``log1p`` / imputation tail, default; "host": numpy), ``stats_engine`` /
    ``predictor_engine`` ("auto": per-gene statistics / correlations on the GPU for large inputs, pandas / numpy otherwise),
    ``device`` (CUDA ordinal) and ``shard`` (a ``parallel.ShardContext``: this process
    trains only its share of the sub-networks and the per-epoch losses / predicted blocks are exchanged with the
    other ranks; see ``deepimpute_b200.parallel``).
    """

    def __init__(self,
                 learning_rate=1e-4,
                 batch_size=64,
                 max_epochs=500,
                 patience=5,
                 ncores=-1,
                 loss="wMSE",
                 output_prefix=None,
                 sub_outputdim=512,
                 verbose=1,
                 seed=1234,
                 architecture=None,
                 *,
                 math_mode=None,
                 device=None,
                 shard=None,
                 predictor_engine="auto",
                 postprocess="gpu",
                 stats_engine="auto",
                 ):
        self.NN_parameters = {"learning_rate": learning_rate,
                              "batch_size": batch_size,
                              "loss": loss,
                              "architecture": architecture,
                              "max_epochs": max_epochs,
                              "patience": patience}
        self.sub_outputdim = sub_outputdim
        # reference default is one mkdtemp() shared by every instance (multinet.py:74); one per instance here
        self.outputdir = output_prefix if output_prefix is not None else tempfile.mkdtemp()
        self.verbose = verbose
        self.seed = seed
        self.setCores(ncores)
        self.math_mode = math_mode
        self.shard = shard
        self.device = device if device is not None else (shard.device if shard is not None else None)
        if predictor_engine not in ("auto", "host", "gpu"):
            raise ValueError("predictor_engine must be 'auto', 'host' or 'gpu'")
        self.predictor_engine = predictor_engine
        if stats_engine not in ("auto", "host", "gpu"):
            raise ValueError("stats_engine must be 'auto', 'host' or 'gpu'")
        # per-gene mean / variance behind both gene filters (multinet.py:191-192, :22-24): pandas reductions on the host
        # (bit-for-bit the reference's numbers) or di_gene_stats on the GPU (large inputs: pandas needs ~45 s at 50k x 20k)
        self.stats_engine = stats_engine
        if postprocess not in ("gpu", "host"):
            raise ValueError("postprocess must be 'gpu' or 'host'")
        # "gpu": log1p of the counts and the whole tail of predict (multinet.py:217, :271, :282-303) run on the device
        # (di_upload_counts / di_impute); "host": the numpy restatement of those lines, which is also what an engine
        # without the fused entry points (the CPU stand-in of the test-suite) is driven with
        self.postprocess = postprocess
        self.timings = {}
        self.engine = None
        self.history = None
        self._owned = None            # sub-networks of every rank (parallel.assign_subnets); None = unsharded

    def _phase(self, name, t0):
        """Wall-clock seconds of one phase of fit / predict, accumulated in ``self.timings`` (bench.py --api)."""
        import time as _time
        now = _time.perf_counter()
        self.timings[name] = self.timings.get(name, 0.0) + (now - t0)
        return now

    def setCores(self, ncores):
        # kept for API compatibility (multinet.py:92-97); the GPU engine ignores it
        if ncores > 0:
            self.ncores = ncores
        else:
            self.ncores = os.cpu_count()
            print("Using all the cores ({})".format(self.ncores))

    def loadDefaultArchitecture(self):
        self.NN_parameters['architecture'] = [
            {"type": "dense", "neurons": self.sub_outputdim // 2, "activation": "relu"},
            {"type": "dropout", "rate": 0.2},
        ]

    # ---- engine seam (reference: build/save/load, multinet.py:105-167) -------------------------------------

    def _parse_architecture(self):
        """Reduce the architecture mini-DSL (``multinet.py:135-143``) to (hidden, dropout_rate)."""
        if self.NN_parameters['architecture'] is None:
            self.loadDefaultArchitecture()
        arch = self.NN_parameters['architecture']
        dense = [l for l in arch if l['type'].lower() == 'dense']
        drop = [l for l in arch if l['type'].lower() == 'dropout']
        other = [l for l in arch if l['type'].lower() not in ('dense', 'dropout')]
        if other:
            print("Unknown layer type.")
        if len(dense) != 1 or len(drop) > 1:
            raise NotImplementedError(
                "the B200 engine implements the hot-path topology Dense(relu) -> Dropout -> Dense(softplus); "
                "got {} dense / {} dropout layers".format(len(dense), len(drop)))
        if dense[0].get('activation', 'relu') != 'relu':
            raise NotImplementedError("hidden activation must be 'relu'")
        if drop and arch.index(drop[0]) < arch.index(dense[0]):
            raise NotImplementedError("dropout must follow the hidden dense layer")
        rate = float(drop[0]['rate']) if drop else 0.0
        return int(dense[0]['neurons']), rate

    def _make_engine(self, inputdims, **kwargs):
        """The one place the CUDA engine is constructed (tests substitute a stand-in here)."""
        from .engine import Engine
        return Engine(inputdims, **kwargs)

    def _my_subnets(self, n_subnets):
        if self._owned is None:
            return list(range(n_subnets))
        return self._owned[self.shard.rank]

    def build(self, inputdims):
        """Create the engine for sub-networks with ``inputdims`` predictors each (reference ``:126-167``).

        Sharded: ``inputdims`` still lists all S sub-networks; this rank's engine holds only the ones
        ``parallel.assign_subnets`` gives it."""
        hidden, rate = self._parse_architecture()
        print(self.NN_parameters['architecture'])
        loss = self.NN_parameters['loss']
        if callable(loss):
            loss = getattr(loss, "__name__", "custom")
        if loss != "wMSE":
            raise NotImplementedError("Unknown loss: {}. The B200 engine fuses wMSE (multinet.py:36-41) only."
                                      .format(loss))
        if self.shard is not None and self.shard.distributed:
            from .parallel import assign_subnets
            self._owned = assign_subnets(inputdims, self.shard.world_size, hidden, self.sub_outputdim)
        else:
            self._owned = None
        mine = self._my_subnets(len(inputdims))
        if not mine:
            raise ValueError("rank {} owns no sub-network: use at most {} ranks".format(self.shard.rank,
                                                                                       len(inputdims)))
        return self._make_engine([inputdims[s] for s in mine],
                                 hidden=hidden,
                                 sub_outputdim=self.sub_outputdim,
                                 learning_rate=self.NN_parameters['learning_rate'],
                                 batch_size=self.NN_parameters['batch_size'],
                                 dropout_rate=rate,
                                 seed=self.seed if self.seed is not None else 0,
                                 math_mode=self.math_mode,
                                 device=self.device,
                                 subnet_ids=mine)

    def _model_path(self):
        if self.shard is None or not self.shard.distributed:
            return "{}/model.npz".format(self.outputdir)
        return "{}/model.rank{}of{}.npz".format(self.outputdir, self.shard.rank, self.shard.world_size)

    def _predict_matrix(self, model, rows=None):
        """[n, S*O] float32 = np.hstack(model.predict(...)) over ALL sub-networks (multinet.py:253, :278-280).

        Sharded: every rank computes its own column block and the blocks are all-gathered (NCCL over NVLink when
        the engine runs on GPUs)."""
        if self._owned is None:
            return model.predict(rows=rows)
        block = model.predict_block(rows=rows)           # torch CUDA tensor (GPU engine) or numpy array
        return self.shard.gather_blocks(block, self._owned, self.sub_outputdim)

    def save(self, model):
        os.makedirs(self.outputdir, exist_ok=True)
        model.save(self._model_path(),
                   targets=self.targets, predictors=[np.asarray(p) for p in self.predictors])
        print("Saved model to disk in {}".format(self.outputdir))

    def load(self):
        from .engine import Engine
        model, extra = Engine.load(self._model_path(), math_mode=self.math_mode, device=self.device)
        if not hasattr(self, "targets"):
            self.targets = extra["targets"]
            self.predictors = [pd.Index(p) for p in extra["predictors"]]
        return model

    # ---- fit (reference multinet.py:169-264) -------------------------------------------------------------

    def fit(self,
            raw,
            cell_subset=1,
            NN_lim=None,
            genes_to_impute=None,
            n_pred=None,
            ntop=5,
            minVMR=0.5,
            mode='random',
            ):
        """Select genes, partition them into sub-networks, train all of them on the GPU, record test metrics."""
        import time as _time
        tp = _time.perf_counter()
        inspect_data(raw)
        tp = self._phase("fit_inspect_s", tp)

        if self.shard is not None and self.shard.distributed and self.seed is None:
            # every rank derives targets / predictors / the cell split from its own np.random stream: without a shared
            # seed the shards would silently train pieces of different models
            raise ValueError("a sharded fit needs a seed (seed=None reseeds every rank from entropy)")
        if self.seed is not None:
            np.random.seed(self.seed)

        if cell_subset != 1:
            raw = raw.sample(frac=cell_subset) if cell_subset < 1 else raw.sample(int(cell_subset))

        cols = raw.columns
        # pandas keeps a homogeneous frame gene-major, so ``raw.values`` is a transposed view: make the cell-major copy the
        # engine wants ONCE and hand the same array to every consumer (statistics, correlations, upload)
        raw_values = np.ascontiguousarray(raw.values)
        tp = self._phase("fit_cell_major_copy_s", tp)
        t0 = _time.perf_counter()
        gpu_stats = self.stats_engine == "gpu" or (self.stats_engine == "auto" and raw_values.size >= 2e8)
        if gpu_stats:
            mean, var, ms = partition.gene_stats_gpu(raw_values, device=self.device if self.device is not None else 0)
            self.timings["gene_stats_device_ms"] = ms
            self._ranked, self._metric = partition.rank_genes_from_stats(mean, var)
        else:
            self._ranked, self._metric = partition.rank_genes(raw)
        if genes_to_impute is None:
            genes = partition.choose_genes(self._ranked, self._metric, self.sub_outputdim, minVMR, NN_lim)
            print("{} genes selected for imputation".format(len(genes)))
        else:
            user = cols.get_indexer(genes_to_impute)
            if (user < 0).any():        # the reference fails here too (KeyError from reindex / .loc, multinet.py:213, :356)
                missing = [g for g, u in zip(genes_to_impute, user) if u < 0]
                raise KeyError("genes_to_impute not in the input columns: {}".format(missing[:10]))
            if len(user) % self.sub_outputdim != 0:
                print("The number of input genes is not a multiple of {}. Filling with other genes."
                      .format(len(user)))
            genes = partition.pad_user_genes(user, self._ranked, self.sub_outputdim)

        if gpu_stats:
            cand = partition.candidate_predictors_from_stats(mean, var, n_pred)
        else:
            cand = partition.candidate_predictors(raw, n_pred)
        self.timings["gene_stats_s"] = _time.perf_counter() - t0
        self.timings["stats_engine"] = "gpu" if gpu_stats else "host"
        # the correlation matrix is O(G^2 N): float64 numpy on the host for small inputs (bit-for-bit the reference's
        # selection), the GPU (fp32, di_corr_topk) for large ones, where the host takes minutes
        # (di_corr_topk keeps at most 8 candidates per target; a wider ntop stays on the host.  With n_pred the device
        # path has the working semantic of DESIGN.md section 9: rows = all targets, columns = the n_pred candidates.)
        use_gpu = ntop <= 8 and (self.predictor_engine == "gpu" or (
            self.predictor_engine == "auto" and float(raw.shape[1]) ** 2 * raw.shape[0] >= 2e11))
        t0 = _time.perf_counter()
        if use_gpu:
            state = np.random.get_state()
            try:
                self._set_partition_gpu(cols, raw_values, genes, cand, ntop, mode)
            except ValueError as exc:
                # a sub-network whose targets cover every candidate: the host path warns and uses all candidates
                # (multinet.py:351-354); take it, from the same point of the random stream
                warnings.warn("GPU predictor selection not applicable ({}); using the host path".format(exc))
                np.random.set_state(state)
                use_gpu = False
        if not use_gpu:
            corr = partition.abs_correlation(raw_values, cand)
            self._set_partition(cols, raw_values, genes, cand, corr, ntop, mode)
        self.timings["predictor_selection_s"] = _time.perf_counter() - t0
        self.timings["predictor_engine"] = "gpu" if use_gpu else "host"

        tp = _time.perf_counter()
        print("Normalization")
        on_device = self.postprocess == "gpu"
        if not on_device:
            norm_values = np.ascontiguousarray(np.log1p(raw_values), dtype=np.float32)

        np.random.seed(self.seed)
        train_rows, test_rows = partition.split_cells(raw.shape[0], labels=_labels(raw.index))
        self.train_cells, self.test_cells = _labels(raw.index)[train_rows], _labels(raw.index)[test_rows]

        print("Building network")
        model = self.build([len(p) for p in self.predictors])

        # The reference materialises 4*S pandas gathers here (multinet.py:231-235); the engine takes the
        # matrix once plus integer index tables and gathers on the device.
        pred_idx = [cols.get_indexer(p).astype(np.int32) for p in self.predictors]
        targ_idx = cols.get_indexer(self.targets.reshape(-1)).reshape(self.targets.shape).astype(np.int32)
        mine = self._my_subnets(len(pred_idx))

        tp = self._phase("fit_build_engine_s", tp)
        print("Fitting with {} cells".format(raw.shape[0]))
        if on_device:       # raw counts go up once; log1p -> float32 (multinet.py:217) happens in HBM
            model.set_counts(raw_values, [pred_idx[s] for s in mine], targ_idx[mine])
        else:
            model.set_data(norm_values, [pred_idx[s] for s in mine], targ_idx[mine])
        tp = self._phase("fit_upload_s", tp)
        # Keras sums the per-branch losses and EarlyStopping watches the sum (multinet.py:242-243): when the
        # branches live on several GPUs the two scalars are summed over ranks before the stop decision
        exchange = None
        if self._owned is not None:
            exchange = lambda epoch, loss, val: self.shard.sum_scalars(loss, val)   # noqa: E731
        result = model.fit(train_rows, test_rows,
                           epochs=self.NN_parameters["max_epochs"],
                           patience=self.NN_parameters["patience"],
                           verbose=self.verbose if (self.shard is None or self.shard.rank == 0) else 0,
                           on_epoch_end=exchange)
        self.history = result.history
        self.trained_epochs = len(result.history['loss'])
        print("Stopped fitting after {} epochs".format(self.trained_epochs))
        tp = self._phase("fit_epochs_s", tp)
        self.timings["fit_epochs_device_ms"] = float(sum(getattr(result, "epoch_ms", []) or []))

        self.engine = model
        self.save(model)
        tp = self._phase("fit_save_s", tp)

        # held-out metrics on originally non-zero entries (multinet.py:251-262)
        if on_device:       # only the held-out block is normalised on the host
            y_true = np.log1p(raw_values[np.ix_(test_rows, targ_idx.reshape(-1))]).astype(np.float32).reshape(-1)
        else:
            y_true = norm_values[np.ix_(test_rows, targ_idx.reshape(-1))].reshape(-1)
        y_hat = self._predict_matrix(model, test_rows).reshape(-1)
        seen = y_true > 0
        y_true, y_hat = y_true[seen], y_hat[seen]
        self.test_metrics = {
            'correlation': pearsonr(y_true, y_hat)[0],
            'MSE': np.sum((y_true - y_hat) ** 2) / len(y_true)
        }
        self._phase("fit_test_metrics_s", tp)
        return self

    def _set_partition(self, cols, raw_values, genes, cand, corr, ntop, mode):
        """targets / predictors as labels, from positional partitioning (multinet.py:213-214)."""
        targ_pos = partition.assign_targets(genes, self.sub_outputdim, mode)
        where = np.full(len(cols), -1, dtype=np.int64)
        where[cand] = np.arange(len(cand))

        def corr_rows(_, t):
            if (where[t] >= 0).all():
                return corr[where[t]]
            # only reachable with n_pred (the reference raises KeyError here, multinet.py:356-358):
            # rows = all targets, columns = the n_pred candidates
            return partition.abs_correlation(raw_values, t, cand)

        pred_pos = partition.choose_predictors(targ_pos, cand, _labels(cols)[cand], corr_rows, ntop)
        self.targets = _labels(cols)[targ_pos]
        self.predictors = [cols[p] for p in pred_pos]

    def _set_partition_gpu(self, cols, raw_values, genes, cand, ntop, mode):
        """Same as ``_set_partition`` with correlations and top-``ntop`` scans on the GPU (``di_corr_topk``)."""
        targ_pos = partition.assign_targets(genes, self.sub_outputdim, mode)
        pred_pos, ms = partition.choose_predictors_gpu(raw_values, targ_pos, cand, _labels(cols)[cand], ntop,
                                                       device=self.device if self.device is not None else 0)
        self.timings["predictor_selection_device_ms"] = ms
        self.targets = _labels(cols)[targ_pos]
        self.predictors = [cols[p] for p in pred_pos]

    # ---- predict (reference multinet.py:266-310) ---------------------------------------------------------

    def predict(self,
                raw,
                imputed_only=False,
                policy="restore"):

        import time as _time
        tp = _time.perf_counter()
        model = self.engine if self.engine is not None else self.load()

        cols = raw.columns
        pred_idx = [cols.get_indexer(p).astype(np.int32) for p in self.predictors]
        targets_flat = self.targets.flatten()
        targ_pos = cols.get_indexer(targets_flat)
        targ_idx = targ_pos.reshape(self.targets.shape).astype(np.int32)
        if self.shard is not None and self.shard.distributed and self._owned is None:
            from .parallel import assign_subnets
            self._owned = assign_subnets([len(p) for p in pred_idx], self.shard.world_size, model.H, model.O)
        mine = self._my_subnets(len(pred_idx))

        if self.postprocess == "gpu":
            # fused path: counts up once, log1p + forward + duplicate mean + clamp + expm1 + policy on the device,
            # one float64 [N, G] matrix back (multinet.py:271-303)
            model.set_counts(raw.values, [pred_idx[s] for s in mine], targ_idx[mine])
            tp = self._phase("predict_upload_s", tp)
            if policy == "restore":
                print("Filling zeros")
            elif policy == "max":
                print("Imputing data with 'max' policy")
            if self._owned is None:
                values = model.impute(policy=policy)
            else:   # sharded: all-gather the prediction blocks on the devices, every rank finishes its own copy
                full = self.shard.gather_blocks_device(model.predict_block(), self._owned, self.sub_outputdim)
                values = model.impute(policy=policy, pred=full, slot_gene=targ_pos)
            # the matrix is freshly allocated and owned by this call: wrap it, do not copy it (pandas would otherwise
            # duplicate all N x G float64 values: ~2 s per GB)
            tp = self._phase("predict_impute_s", tp)
            imputed = pd.DataFrame(values, index=raw.index, columns=raw.columns, copy=False)
            self._phase("predict_wrap_s", tp)
            if imputed_only:
                return imputed.loc[:, np.unique(targets_flat)]
            return imputed

        norm_raw = np.log1p(raw)
        norm_values = norm_raw.values
        model.set_data(np.ascontiguousarray(norm_values, dtype=np.float32), [pred_idx[s] for s in mine],
                       targ_idx[mine])

        predicted = self._predict_matrix(model)           # [N, S*O] float32, columns = targets.flatten()

        # mean over duplicated target columns (multinet.py:284): the reference's own pandas operation -- a column
        # groupby-mean of a float32 frame -- spelled through a transpose because pandas 3 dropped groupby(axis=1)
        uniq_labels = np.unique(targets_flat)
        uniq_pos = cols.get_indexer(uniq_labels)
        grouped = pd.DataFrame(predicted, columns=targets_flat).T.groupby(level=0).mean().T
        pred_u = grouped.loc[:, uniq_labels].values

        imputed = np.array(norm_values, dtype=np.float64, copy=True)
        imputed[:, uniq_pos] = pred_u

        # To prevent overflow
        imputed[(imputed > 2 * norm_values.max()) | (np.isnan(imputed))] = 0
        # Convert back to counts
        imputed = np.expm1(imputed)

        if policy == "restore":
            print("Filling zeros")
            mask = (raw.values > 0)
            imputed[mask] = raw.values[mask]
        elif policy == "max":
            print("Imputing data with 'max' policy")
            mask = (raw.values > imputed)
            imputed[mask] = raw.values[mask]

        imputed = pd.DataFrame(imputed, index=raw.index, columns=raw.columns, copy=False)

        if imputed_only:
            return imputed.loc[:, uniq_labels]
        else:
            return imputed

    # ---- reference-named entry points for the partitioning steps (multinet.py:312-365) ------------------

    def filter_genes(self, gene_metric, threshold, NN_lim=None):
        """``gene_metric``: Series sorted descending, as in the reference; returns gene labels."""
        pos = partition.choose_genes(np.arange(len(gene_metric)), gene_metric.values,
                                     self.sub_outputdim, threshold, NN_lim)
        print("{} genes selected for imputation".format(len(pos)))
        return _labels(gene_metric.index)[pos]

    def setTargets(self, data, mode='random'):
        pos = partition.assign_targets(np.arange(data.shape[1]), self.sub_outputdim, mode)
        self.targets = _labels(data.columns)[pos]

    def setPredictors(self, covariance_matrix, ntop=5):
        labels = covariance_matrix.columns
        values = covariance_matrix.values
        row_of = pd.Series(np.arange(len(covariance_matrix.index)), index=covariance_matrix.index)
        targ_pos = [labels.get_indexer(t) for t in self.targets]          # -1 = target is no candidate

        def corr_rows(i, _):
            return values[row_of.loc[self.targets[i]].values]

        pred = partition.choose_predictors(targ_pos, np.arange(len(labels)), _labels(labels), corr_rows, ntop)
        self.predictors = [labels[p] for p in pred]

    def score(self, data, policy=None):
        warnings.warn(
            "This method is deprecated. Please use model.test_metrics to measure model accuracy instead",
            DeprecationWarning)
        Y_hat = self.predict(data, policy=policy)
        Y = data.loc[Y_hat.index, Y_hat.columns]

        return pearsonr(Y_hat.values.reshape(-1), Y.values.reshape(-1))
