"""CPU restatement of the element-wise ends of the path -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``log1p_norm``   reference ``multinet.py:217`` (``np.log1p(raw).astype(np.float32)``), ``:271``
``impute_tail``  reference ``multinet.py:282-303``: duplicate-column mean, un-imputed genes carried as float64
                 ``log1p(raw)``, overflow clamp, ``expm1``, then the ``restore`` / ``max`` policy

PINNED: ``tests/golden/predict_tail.npz`` holds the output of the reference's own, unmodified ``MultiNet.predict`` for a
fixed prediction matrix (``scripts/make_golden_tail.py``: the Keras model behind ``self.load()`` is a stand-in that returns
that matrix, and ``groupby(axis=1)`` -- removed in pandas 3 -- is routed to its transpose spelling);
``tests/test_postprocess.py`` checks this restatement bit-for-bit against those vectors and, in the build container,
against the reference running live on further seeded cases.  The restatement keeps the reference's own pandas
operations with gene POSITIONS as labels.  Note ``:284`` on a float32 frame is pandas' float32 group mean: NaN entries
skipped, Kahan summation in float32, division by the count in float32 -- the CUDA kernel follows exactly that.
"""
import numpy as np
import pandas as pd


def log1p_norm(raw):
    """``np.log1p(raw).astype(np.float32)`` as the reference evaluates it on its float64 frames (``pd.read_csv`` gives
    float64 / int64 columns, so numpy's float64 ``log1p`` runs and the result is rounded once).  A float32 frame would
    send numpy down ``log1pf`` (not correctly rounded, up to 1 float32 ulp away); the device always computes the
    float64 form, so that is what the oracle states for every input dtype."""
    return np.log1p(np.asarray(raw, dtype=np.float64)).astype(np.float32)


def impute_tail(raw, predicted, slot_gene, policy="restore"):
    """raw [N, G] counts; predicted [N, n_slots] float32, column k predicts gene ``slot_gene[k]``.  Returns float64 [N, G]."""
    # float64 frame, as pd.read_csv delivers it (a float32 frame would make numpy carry log1pf results in float32 through
    # :271-:293; the device states the float64 form for every input dtype, see log1p_norm)
    raw = pd.DataFrame(np.asarray(raw, dtype=np.float64))
    norm_raw = np.log1p(raw)                                                    # :271
    predicted = pd.DataFrame(np.asarray(predicted, dtype=np.float32), columns=np.asarray(slot_gene))
    predicted = predicted.T.groupby(level=0).mean().T                           # :284 (axis=1 groupby, pandas-3 spelling)
    not_predicted = norm_raw.drop(predicted.columns, axis=1)                    # :286
    imputed = pd.concat([predicted, not_predicted], axis=1).loc[raw.index, raw.columns].values   # :287
    imputed = imputed.astype(np.float64)
    imputed[(imputed > 2 * norm_raw.values.max()) | (np.isnan(imputed))] = 0    # :291
    imputed = np.expm1(imputed)                                                 # :293
    if policy == "restore":                                                     # :295-298
        mask = raw.values > 0
        imputed[mask] = raw.values[mask]
    elif policy == "max":                                                       # :299-302
        mask = raw.values > imputed
        imputed[mask] = raw.values[mask]
    return imputed
