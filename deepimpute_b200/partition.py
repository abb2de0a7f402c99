"""Gene selection, output-gene partitioning and predictor selection in positional (integer) index space.

These are the host-side steps of the reference that decide *what* the engine trains: how many sub-networks
there are, which 512 genes each one outputs and which genes feed it.  They must give the same answer as the
reference for the same ``np.random`` seed, so every draw from the legacy global stream happens here in the
reference's order and with calls that consume the stream identically:

=====================  ===============================  =============================================
step                   reference                        draw
=====================  ===============================  =============================================
filler genes           ``multinet.py:326`` / ``:207``   ``choice(n, rest)``  (with replacement)
target partition       ``multinet.py:340-342``          ``choice(n, [S, O], replace=False)``
held-out cells         ``multinet.py:228``              ``choice(N, int(0.05*N), replace=False)``
=====================  ===============================  =============================================

``np.random.choice(labels, ...)`` draws integer positions and then indexes ``labels``; drawing positions
directly is the same stream, so everything below works on positions and the caller maps back to labels.
"""
import warnings

import numpy as np
import pandas as pd


def rank_genes(raw):
    """Positions of genes with var/(1+mean) > 0, best first (reference ``multinet.py:191-192``).

    pandas reductions and ``sort_values`` are used on purpose: their summation order and tie-breaking are what
    the reference sees.
    """
    metric = raw.var() / (1 + raw.mean())
    metric = pd.Series(metric.values, index=np.arange(raw.shape[1])).sort_values(ascending=False)
    metric = metric[metric > 0]
    return metric.index.values.astype(np.int64), metric.values


def gene_stats_gpu(raw_values, device=0):
    """Per-gene (mean, variance with ddof = 1) of the count matrix on the GPU (``di_gene_stats``), float64.
    Returns (mean, var, device milliseconds)."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    x = np.asarray(raw_values)
    if x.dtype != np.float32:
        x = x.astype(np.float64, copy=False)
    x = np.ascontiguousarray(x)
    mean, var = np.empty(x.shape[1], np.float64), np.empty(x.shape[1], np.float64)
    ms = C.c_float()
    dp = C.POINTER(C.c_double)
    rc = lib.di_gene_stats(int(device), C.c_void_p(x.ctypes.data), _lib.DI_DTYPE[x.dtype.name], x.shape[0], x.shape[1],
                           mean.ctypes.data_as(dp), var.ctypes.data_as(dp), C.byref(ms))
    if rc != 0:
        raise RuntimeError("di_gene_stats failed ({}): {}".format(rc, lib.di_gene_stats_last_error().decode()))
    return mean, var, ms.value


def rank_genes_from_stats(mean, var):
    """``rank_genes`` from precomputed per-gene statistics (same ranking rule, ``multinet.py:191-192``)."""
    metric = pd.Series(var / (1 + mean), index=np.arange(len(mean))).sort_values(ascending=False)
    metric = metric[metric > 0]
    return metric.index.values.astype(np.int64), metric.values


def candidate_predictors_from_stats(mean, var, n_pred=None):
    """``candidate_predictors`` from precomputed per-gene statistics (``multinet.py:22-29``)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        cv = np.sqrt(var) / mean
    cv[np.isinf(cv)] = 0
    cv = pd.Series(cv, index=np.arange(len(mean)))
    if n_pred is None:
        return np.flatnonzero((cv > 0).values)
    print("Using {} predictors".format(n_pred))
    return cv.sort_values(ascending=False).index.values[:n_pred].astype(np.int64)


def choose_genes(ranked, metric_values, sub_outputdim, threshold, limit=None):
    """Genes to impute, padded with random filler genes to a multiple of ``sub_outputdim``.

    Follows ``filter_genes`` (``multinet.py:312-331``), including its quirk: the pad length is
    ``O - (len % O)``, which is a whole extra sub-network of random genes when ``len % O == 0``.
    """
    if not str(limit).isdigit():
        limit = int((metric_values > threshold).sum())
    n_nets = int(np.ceil(int(limit) / sub_outputdim))
    chosen = ranked[:n_nets * sub_outputdim]
    pad = sub_outputdim - (len(chosen) % sub_outputdim)
    if pad > 0:
        chosen = np.concatenate([chosen, ranked[np.random.choice(len(ranked), pad)]])
    return chosen


def pad_user_genes(user_genes, ranked, sub_outputdim):
    """User-supplied gene list padded up to a whole number of sub-networks (``multinet.py:196-209``).

    For fewer than ``sub_outputdim`` genes this is the reference's rule: the best-ranked genes fill the sub-network,
    random ones (with replacement) top it up when the ranking is too short.  For MORE than ``sub_outputdim`` genes
    that are not a multiple of it the reference's slice ``gene_metric.index[:O - n]`` goes negative and appends almost
    every gene; here the list is padded to the next multiple of ``sub_outputdim`` by the same rule, so every requested
    gene is imputed and nothing else changes (deliberate difference)."""
    user_genes = np.asarray(user_genes)
    n = len(user_genes)
    need = (-n) % sub_outputdim
    if need == 0:
        return user_genes
    filler = ranked[:need]
    short = need - len(filler)
    if short > 0:
        filler = np.concatenate([filler, ranked[np.random.choice(len(ranked), short, replace=True)]])
    return np.concatenate([user_genes, filler])


def assign_targets(genes, sub_outputdim, mode="random"):
    """[S, O] target genes per sub-network (``setTargets``, ``multinet.py:333-342``)."""
    n_nets = int(len(genes) / sub_outputdim)
    if mode == "progressive":
        return np.asarray(genes)[:n_nets * sub_outputdim].reshape(n_nets, sub_outputdim)
    pick = np.random.choice(len(genes), [n_nets, sub_outputdim], replace=False)
    return np.asarray(genes)[pick]


def candidate_predictors(raw, n_pred=None):
    """Positions of candidate predictor genes: std/mean > 0, or the top ``n_pred`` of it (``multinet.py:22-29``)."""
    cv = raw.std() / raw.mean()
    cv[np.isinf(cv)] = 0
    cv = pd.Series(cv.values, index=np.arange(raw.shape[1]))
    if n_pred is None:
        return np.flatnonzero((cv > 0).values)
    print("Using {} predictors".format(n_pred))
    return cv.sort_values(ascending=False).index.values[:n_pred].astype(np.int64)


def abs_correlation(raw_values, rows, cols=None):
    """|corrcoef| of gene columns ``rows`` x ``cols`` on raw counts, NaN -> 0 (``multinet.py:31-33``).

    With ``cols is None`` this is exactly ``np.abs(np.corrcoef(raw.T.loc[rows]))``.
    """
    if cols is None:
        c = np.abs(np.corrcoef(raw_values[:, rows].T))
        return np.nan_to_num(c, nan=0.0, posinf=np.inf, neginf=-np.inf)
    # only the columns involved are widened to float64 (never the whole matrix: 48 GB at 200k x 30k)
    a = np.asarray(raw_values[:, rows], dtype=np.float64)
    b = np.asarray(raw_values[:, cols], dtype=np.float64)
    a -= a.mean(0)
    b -= b.mean(0)
    with np.errstate(invalid="ignore", divide="ignore"):
        c = (a.T @ b) / np.sqrt(np.outer((a * a).sum(0), (b * b).sum(0)))
    return np.nan_to_num(np.abs(np.clip(c, -1, 1)), nan=0.0)


def choose_predictors(targets, cand, cand_labels, corr_rows, ntop=5):
    """Predictor genes of every sub-network (``setPredictors``, ``multinet.py:344-365``).

    ``targets``      [S, O] gene positions.
    ``cand``         candidate gene positions, in correlation-matrix column order.
    ``cand_labels``  their labels: the reference takes ``np.setdiff1d`` on *labels*, so the surviving
                     candidates are visited in label-sorted order and ties in the top-``ntop`` break that way.
    ``corr_rows``    callable ``(subnet, target_positions) -> [len, len(cand)]`` |r| rows.

    Per target: the ``ntop`` most correlated candidates outside the sub-network's own targets; per sub-network:
    their union in first-appearance order.
    """
    cand = np.asarray(cand)
    label_order = np.argsort(np.asarray(cand_labels), kind="stable")      # np.setdiff1d returns sorted labels
    out = []
    for i, t in enumerate(targets):
        keep = label_order[~np.isin(cand[label_order], t)]
        if keep.size == 0:
            warnings.warn('Warning: number of target genes lower than output dim. Consider lowering down the '
                          'sub_outputdim parameter', UserWarning)
            keep = np.arange(len(cand))
        sub = corr_rows(i, t)[:, keep]
        top = np.argsort(-sub, axis=1)[:, :ntop].ravel()
        picked = pd.unique(cand[keep][top])
        out.append(picked.astype(np.int64))
        print("Net {}: {} predictors, {} targets".format(i, len(picked), len(t)))
    return out


def choose_predictors_gpu(raw_values, targets, cand, cand_labels, ntop=5, device=0):
    """``choose_predictors`` with the correlation matrix and the per-target top-``ntop`` scan on the GPU
    (``di_corr_topk``): same selection rule, fp32 instead of float64 correlations (differences only at ties closer
    than ~1e-6).  Returns (predictor positions per sub-network, device milliseconds)."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    cand = np.asarray(cand)
    order = np.argsort(np.asarray(cand_labels), kind="stable")            # the reference's label-sorted visiting order
    cand_sorted = np.ascontiguousarray(cand[order], dtype=np.int32)
    targets = np.ascontiguousarray(targets, dtype=np.int32)
    n_nets, out_dim = targets.shape
    raw32 = np.ascontiguousarray(raw_values, dtype=np.float32)
    top = np.empty((n_nets, out_dim, ntop), dtype=np.int32)
    ms = C.c_float()
    rc = lib.di_corr_topk(int(device), _lib.f32(raw32), raw32.shape[0], raw32.shape[1], _lib.i32(cand_sorted),
                          len(cand_sorted), _lib.i32(targets), n_nets, out_dim, int(ntop), _lib.i32(top), None,
                          C.byref(ms))
    if rc != 0:
        raise RuntimeError("di_corr_topk failed ({}): {}".format(rc, lib.di_corr_last_error().decode()))
    out = []
    for i in range(n_nets):
        flat = top[i].reshape(-1)
        if (flat < 0).any():
            raise ValueError("sub-network {} has fewer than {} candidate predictors outside its own targets".format(i, ntop))
        picked = pd.unique(cand_sorted[flat])
        out.append(picked.astype(np.int64))
        print("Net {}: {} predictors, {} targets".format(i, len(picked), out_dim))
    return out, ms.value


def split_cells(n_cells, labels=None):
    """5 % held-out cells (``multinet.py:228-229``): returns (train_rows, test_rows).

    Test rows keep draw order; train rows are sorted by *label* because the reference builds them with
    ``np.setdiff1d`` on the index labels.
    """
    test = np.random.choice(n_cells, int(0.05 * n_cells), replace=False)
    mask = np.ones(n_cells, dtype=bool)
    mask[test] = False
    train = np.flatnonzero(mask)
    if labels is not None:
        train = train[np.argsort(np.asarray(labels)[train], kind="stable")]
    return train.astype(np.int32), test.astype(np.int32)
