"""``Engine``: the object that takes the place of the Keras ``Model`` in the reference.

The reference drives its network through four verbs of the Keras functional ``Model`` --
``fit(X_list, Y_list, validation_data, epochs, batch_size, callbacks=[EarlyStopping])`` (multinet.py:238-244),
``predict(X_list)`` (:253, :278), ``save_weights``/``to_json`` (:108-114) and ``load_weights`` (:121-122).
``Engine`` offers the same verbs over the C-ABI of ``include/deepimpute_b200.h``.  Its native input is the
normalised matrix plus index tables (``set_data`` / ``fit`` / ``predict``), so the per-sub-network gathers of
multinet.py:231-235 happen on the GPU; ``fit_arrays`` / ``predict_arrays`` take the reference's lists of arrays.

Host-side policy that Keras implements in Python stays in Python here: Glorot initialisation, the per-epoch
shuffle and the EarlyStopping rule.  All arithmetic is in the CUDA library; nothing here computes on the CPU.
"""
import ctypes as C
import os

import numpy as np

from . import _lib


DEFAULT_MATH = "tf32x3"


class History:
    """Stand-in for ``keras.callbacks.History``: ``.history['loss']`` / ``['val_loss']`` (multinet.py:246)."""

    def __init__(self):
        self.history = {"loss": [], "val_loss": []}
        self.epoch_ms = []


def glorot_uniform(n_pred, hidden, out, seed, subnet_ids=None):
    """Keras ``Dense`` defaults: kernel ~ U(+-sqrt(6/(fan_in+fan_out))), zero bias; one stream per sub-network,
    keyed by its global number so that a sharded model starts from the same weights as an unsharded one."""
    weights = []
    ids = range(len(n_pred)) if subnet_ids is None else subnet_ids
    for s, p in zip(ids, n_pred):
        rng = np.random.default_rng([int(seed), int(s)])
        a1 = np.sqrt(6.0 / (p + hidden))
        a2 = np.sqrt(6.0 / (hidden + out))
        W1 = rng.uniform(-a1, a1, size=(p, hidden)).astype(np.float32)
        W2 = rng.uniform(-a2, a2, size=(hidden, out)).astype(np.float32)
        weights.append((W1, np.zeros(hidden, np.float32), W2, np.zeros(out, np.float32)))
    return weights


def epoch_permutation(seed, epoch, n):
    """Order in which epoch ``epoch`` visits the n training rows (Keras ``fit(shuffle=True)``)."""
    return np.random.default_rng([int(seed), 0x5EED, int(epoch)]).permutation(n).astype(np.int32)


class PermutationPrefetcher:
    """Visiting orders one epoch ahead of the GPU.

    Drawing the permutation of ~50k training rows costs about a millisecond of host time; ``di_train_epoch`` blocks
    until the epoch is done and ctypes releases the GIL meanwhile, so a helper thread draws the order of epoch e + 1
    while the device runs epoch e.  ``perm_fn(epoch)`` must be a pure function of the epoch number (it is:
    ``epoch_permutation``), so which thread evaluates it -- and an order drawn for an epoch that early stopping never
    runs -- changes nothing.
    """

    def __init__(self, perm_fn, n_epochs=None):
        from concurrent.futures import ThreadPoolExecutor
        self._fn = perm_fn
        self._end = n_epochs                    # never ask perm_fn for an epoch at or beyond this one
        self._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="di-perm")
        self._next = None                       # (epoch, future)

    def get(self, epoch):
        nxt, self._next = self._next, None
        perm = nxt[1].result() if (nxt is not None and nxt[0] == epoch) else self._fn(epoch)
        if self._end is None or epoch + 1 < self._end:
            self._next = (epoch + 1, self._pool.submit(self._fn, epoch + 1))
        return perm

    def close(self):
        self._next = None
        self._pool.shutdown(wait=False, cancel_futures=True)


class Engine:
    def __init__(self, inputdims, hidden=256, sub_outputdim=512, learning_rate=1e-4, batch_size=64,
                 dropout_rate=0.2, seed=1234, beta1=0.9, beta2=0.999, epsilon=1e-7,
                 math_mode=None, device=None, init_weights=True, subnet_ids=None):
        self._h = C.c_void_p()
        self.lib = _lib.load()
        self.n_pred = [int(p) for p in inputdims]
        self.S, self.H, self.O = len(self.n_pred), int(hidden), int(sub_outputdim)
        self.B, self.seed = int(batch_size), int(seed)
        self.learning_rate, self.dropout_rate = float(learning_rate), float(dropout_rate)
        self.beta1, self.beta2, self.epsilon = float(beta1), float(beta2), float(epsilon)
        math_mode = math_mode or os.environ.get("DEEPIMPUTE_B200_MATH", DEFAULT_MATH)
        if math_mode not in _lib.DI_MATH:
            raise ValueError("math_mode must be one of {}".format(sorted(_lib.DI_MATH)))
        self.math_mode = math_mode
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = int(device)
        cfg = _lib.DiConfig(self.S, self.H, self.O, self.B, self.learning_rate, self.beta1, self.beta2,
                            self.epsilon, self.dropout_rate, self.seed & (2 ** 64 - 1),
                            _lib.DI_MATH[math_mode], self.device)
        n_pred = np.asarray(self.n_pred, dtype=np.int32)
        rc = self.lib.di_create(C.byref(self._h), C.byref(cfg), _lib.i32(n_pred))
        if rc != 0:
            msg = self.lib.di_last_error(None)
            self._h = C.c_void_p()
            raise RuntimeError("di_create failed ({}): {}".format(rc, msg.decode() if msg else "?"))
        self.subnet_ids = list(range(self.S)) if subnet_ids is None else [int(s) for s in subnet_ids]
        if len(self.subnet_ids) != self.S:
            raise ValueError("subnet_ids must name every sub-network")
        if subnet_ids is not None:
            ids = np.asarray(self.subnet_ids, dtype=np.int32)
            self._check(self.lib.di_set_subnet_ids(self._h, _lib.i32(ids)))
        self.n_cells = self.n_genes = None
        self.has_counts = False
        self.n_train = self.n_test = 0
        self.steps_done = 0
        if init_weights:
            self.set_weights(glorot_uniform(self.n_pred, self.H, self.O, self.seed, self.subnet_ids))

    # -- plumbing -------------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            msg = self.lib.di_last_error(self._h)
            raise RuntimeError("deepimpute_b200 error {}: {}".format(rc, msg.decode() if msg else "?"))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.di_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data -----------------------------------------------------------------------------------------------
    def _check_partition(self, n_genes, pred_idx, targ_idx):
        if len(pred_idx) != self.S or [len(p) for p in pred_idx] != self.n_pred:
            raise ValueError("pred_idx does not match the engine's input dims")
        targ_idx = np.ascontiguousarray(targ_idx, dtype=np.int32)
        if targ_idx.shape != (self.S, self.O):
            raise ValueError("targ_idx must be [S, O]")
        flat = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.int32) for p in pred_idx]))
        off = np.concatenate([[0], np.cumsum(self.n_pred)]).astype(np.int64)
        for name, idx in (("pred_idx", flat), ("targ_idx", targ_idx)):
            if idx.size and (idx.min() < 0 or idx.max() >= n_genes):
                raise ValueError("{} out of range for a matrix with {} genes".format(name, n_genes))
        return flat, off, targ_idx

    def set_data(self, norm, pred_idx, targ_idx):
        """norm [N, G] float32 (log1p of counts); pred_idx: S int arrays of gene columns; targ_idx [S, O]."""
        norm = np.ascontiguousarray(norm, dtype=np.float32)
        if norm.ndim != 2:
            raise ValueError("norm must be 2-D")
        flat, off, targ_idx = self._check_partition(norm.shape[1], pred_idx, targ_idx)
        self._check(self.lib.di_upload_matrix(self._h, _lib.f32(norm), norm.shape[0], norm.shape[1]))
        self._check(self.lib.di_set_partition(self._h, _lib.i32(flat), _lib.i64(off), _lib.i32(targ_idx)))
        self.n_cells, self.n_genes = norm.shape
        self.has_counts = False
        self.n_train = self.n_test = 0

    def set_counts(self, raw, pred_idx, targ_idx):
        """Like ``set_data`` but from RAW counts [N, G] (float32 or float64): ``log1p`` -> float32 runs on the device
        (multinet.py:217, :271) and the counts stay resident for ``impute``."""
        raw = np.asarray(raw)
        if raw.dtype != np.float32:
            raw = raw.astype(np.float64, copy=False)
        raw = np.ascontiguousarray(raw)
        if raw.ndim != 2:
            raise ValueError("raw must be 2-D")
        flat, off, targ_idx = self._check_partition(raw.shape[1], pred_idx, targ_idx)
        self._check(self.lib.di_upload_counts(self._h, C.c_void_p(raw.ctypes.data), _lib.DI_DTYPE[raw.dtype.name],
                                              raw.shape[0], raw.shape[1]))
        self._check(self.lib.di_set_partition(self._h, _lib.i32(flat), _lib.i64(off), _lib.i32(targ_idx)))
        self.n_cells, self.n_genes = raw.shape
        self.has_counts = True
        self.n_train = self.n_test = 0

    def set_split(self, train_rows, test_rows):
        tr = np.ascontiguousarray(train_rows, dtype=np.int32)
        te = np.ascontiguousarray(test_rows, dtype=np.int32)
        for r in (tr, te):
            if r.size and (r.min() < 0 or r.max() >= self.n_cells):
                raise ValueError("row index out of range")
        self._check(self.lib.di_set_split(self._h, _lib.i32(tr), len(tr), _lib.i32(te), len(te)))
        self.n_train, self.n_test = len(tr), len(te)

    # -- weights --------------------------------------------------------------------------------------------
    def set_weights(self, weights):
        if len(weights) != self.S:
            raise ValueError("expected weights for {} sub-networks".format(self.S))
        for s, (W1, b1, W2, b2) in enumerate(weights):
            W1, b1, W2, b2 = [np.ascontiguousarray(a, dtype=np.float32) for a in (W1, b1, W2, b2)]
            if W1.shape != (self.n_pred[s], self.H) or b1.shape != (self.H,) or \
                    W2.shape != (self.H, self.O) or b2.shape != (self.O,):
                raise ValueError("bad weight shapes for sub-network {}".format(s))
            self._check(self.lib.di_set_weights(self._h, s, _lib.f32(W1), _lib.f32(b1), _lib.f32(W2), _lib.f32(b2)))
        self.steps_done = 0

    def get_weights(self):
        out = []
        for s in range(self.S):
            W1 = np.empty((self.n_pred[s], self.H), np.float32)
            b1 = np.empty(self.H, np.float32)
            W2 = np.empty((self.H, self.O), np.float32)
            b2 = np.empty(self.O, np.float32)
            self._check(self.lib.di_get_weights(self._h, s, _lib.f32(W1), _lib.f32(b1), _lib.f32(W2), _lib.f32(b2)))
            out.append((W1, b1, W2, b2))
        return out

    def get_adam_state(self, s):
        shapes = [(self.n_pred[s], self.H), (self.H,), (self.H, self.O), (self.O,)]
        bufs = []
        for shp in shapes:
            bufs += [np.empty(shp, np.float32), np.empty(shp, np.float32)]
        t = C.c_int64()
        self._check(self.lib.di_get_adam_state(self._h, s, *[_lib.f32(b) for b in bufs], C.byref(t)))
        return bufs, t.value

    # -- training -------------------------------------------------------------------------------------------
    def train_step(self, rows, step=None):
        """One Adam step on the given cells (positions into norm, at most batch_size); returns the summed wMSE."""
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        step = self.steps_done if step is None else int(step)
        loss = C.c_float()
        self._check(self.lib.di_train_step(self._h, _lib.i32(rows), len(rows), step, C.byref(loss)))
        self.steps_done = step + 1
        return loss.value

    def train_epoch(self, perm, first_step=None):
        perm = np.ascontiguousarray(perm, dtype=np.int32)
        if len(perm) != self.n_train:
            raise ValueError("perm must permute the {} training rows".format(self.n_train))
        first_step = self.steps_done if first_step is None else int(first_step)
        loss, val = C.c_float(), C.c_float()
        self._check(self.lib.di_train_epoch(self._h, _lib.i32(perm), first_step, C.byref(loss), C.byref(val)))
        self.steps_done = first_step + (self.n_train + self.B - 1) // self.B
        return loss.value, val.value

    def validation_loss(self):
        val = C.c_float()
        self._check(self.lib.di_validation_loss(self._h, C.byref(val)))
        return val.value

    def fit(self, train_rows, test_rows, epochs, patience=5, verbose=1, perm_fn=None, on_epoch_end=None):
        """Keras ``model.fit`` + ``EarlyStopping(monitor='val_loss', patience)`` (multinet.py:238-244).

        ``on_epoch_end(epoch, loss, val_loss) -> (loss, val_loss)`` lets a multi-GPU caller all-reduce the two
        scalars so that every shard takes the same early-stopping decision (deepimpute_b200.parallel).
        """
        self.set_split(train_rows, test_rows)
        hist = History()
        n_train, seed = self.n_train, self.seed
        perms = PermutationPrefetcher(perm_fn or (lambda e: epoch_permutation(seed, e, n_train)), int(epochs))
        try:
            return self._fit_epochs(hist, perms, int(epochs), patience, verbose, on_epoch_end)
        finally:
            perms.close()

    def _fit_epochs(self, hist, perms, epochs, patience, verbose, on_epoch_end):
        best, wait = np.inf, 0
        for epoch in range(epochs):
            loss, val = self.train_epoch(perms.get(epoch))
            hist.epoch_ms.append(self.lib.di_last_device_ms(self._h))
            if on_epoch_end is not None:
                loss, val = on_epoch_end(epoch, loss, val)
            hist.history["loss"].append(loss)
            hist.history["val_loss"].append(val)
            if verbose:
                print("Epoch {}/{} - loss: {:.6f} - val_loss: {:.6f}".format(epoch + 1, epochs, loss, val))
            if not np.isfinite(loss):
                raise FloatingPointError("non-finite training loss at epoch {}".format(epoch + 1))
            if val < best:
                best, wait = val, 0
            else:
                wait += 1
                if wait >= patience:
                    break
        return hist

    # -- inference ------------------------------------------------------------------------------------------
    def predict(self, rows=None, out=None):
        """[n, S*O] float32; column s*O+o is target gene targ_idx[s][o] (np.hstack of the Keras outputs)."""
        if self.n_cells is None:
            raise RuntimeError("set_data() first")
        if rows is None:
            n, rp = self.n_cells, None
        else:
            rows = np.ascontiguousarray(rows, dtype=np.int32)
            n, rp = len(rows), _lib.i32(rows)
        if out is None:
            out = np.empty((n, self.S * self.O), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (n, self.S * self.O)
        self._check(self.lib.di_predict(self._h, rp, n, _lib.f32(out)))
        return out

    def predict_device(self, d_out_ptr, ld_out, rows=None):
        """Write the prediction block into device memory (e.g. a slice of a torch all-gather buffer)."""
        if rows is None:
            n, rp = self.n_cells, None
        else:
            rows = np.ascontiguousarray(rows, dtype=np.int32)
            n, rp = len(rows), _lib.i32(rows)
        self._check(self.lib.di_predict_device(self._h, rp, n, C.c_void_p(int(d_out_ptr)), int(ld_out)))

    def predict_block(self, rows=None):
        """This engine's prediction block as a torch CUDA tensor [n, S*O] (input of the multi-GPU all-gather)."""
        import torch
        n = self.n_cells if rows is None else len(rows)
        out = torch.empty((n, self.S * self.O), dtype=torch.float32, device=torch.device("cuda", self.device))
        self.predict_device(out.data_ptr(), self.S * self.O, rows)
        return out

    def impute(self, policy="restore", out=None, dtype=np.float64, pred=None, slot_gene=None):
        """The fused tail of ``MultiNet.predict`` (multinet.py:278-303) for all cells: forward, duplicate-target mean,
        overflow clamp, ``expm1`` and the ``policy`` against the resident counts -> ``[N, G]`` array of ``dtype``.

        ``pred``: optional torch CUDA tensor ``[N, n_slots]`` float32 of predictions made elsewhere (the all-gathered
        blocks of a sharded model) with ``slot_gene[n_slots]`` naming the gene column each prediction column targets;
        by default the engine predicts with its own sub-networks.  Needs ``set_counts``."""
        if not self.has_counts:
            raise RuntimeError("impute() needs the raw counts on the device: call set_counts() first")
        if policy not in _lib.DI_POLICY:
            policy = "none"          # the reference ignores unknown policies (multinet.py:295-302)
        dtype = np.dtype(dtype)
        if dtype.name not in _lib.DI_DTYPE:
            raise ValueError("dtype must be float32 or float64")
        if out is None:
            out = np.empty((self.n_cells, self.n_genes), dtype=dtype)
        if out.dtype != dtype or not out.flags.c_contiguous or out.shape != (self.n_cells, self.n_genes):
            raise ValueError("out must be a C-contiguous [N, G] array of the requested dtype")
        sg, n_slots, dp, ld = None, 0, None, 0
        if slot_gene is not None:
            slot_gene = np.ascontiguousarray(slot_gene, dtype=np.int32).reshape(-1)
            sg, n_slots = _lib.i32(slot_gene), len(slot_gene)
        if pred is not None:
            if slot_gene is None:
                raise ValueError("pred needs slot_gene")
            if pred.dtype.__str__() != "torch.float32" or not pred.is_cuda or pred.dim() != 2 or \
                    pred.shape[0] != self.n_cells or pred.shape[1] < n_slots or pred.stride(1) != 1:
                raise ValueError("pred must be a float32 CUDA tensor [N, >= n_slots] with unit column stride")
            import torch
            torch.cuda.current_stream(pred.device).synchronize()
            dp, ld = C.c_void_p(pred.data_ptr()), pred.stride(0)
        self._check(self.lib.di_impute(self._h, _lib.DI_POLICY[policy], sg, n_slots, dp, ld,
                                       _lib.DI_DTYPE[dtype.name], C.c_void_p(out.ctypes.data)))
        return out

    # -- Keras-shaped adapters (lists of arrays, as the reference passes them) -----------------------------
    def _concat(self, X_list, Y_list=None):
        n = X_list[0].shape[0]
        blocks = [np.asarray(x, dtype=np.float32) for x in X_list]
        pred_idx, c = [], 0
        for x in blocks:
            pred_idx.append(np.arange(c, c + x.shape[1], dtype=np.int32))
            c += x.shape[1]
        if Y_list is not None:
            blocks += [np.asarray(y, dtype=np.float32) for y in Y_list]
        else:
            blocks.append(np.zeros((n, self.S * self.O), np.float32))
        targ_idx = np.arange(c, c + self.S * self.O, dtype=np.int32).reshape(self.S, self.O)
        return np.concatenate(blocks, axis=1), pred_idx, targ_idx

    def fit_arrays(self, X_list, Y_list, validation_data, epochs, patience=5, verbose=0, perm_fn=None):
        """``model.fit(X_train, Y_train, validation_data=(X_test, Y_test), ...)`` with the reference's arguments."""
        Xv, Yv = validation_data
        tr, pred_idx, targ_idx = self._concat(X_list, Y_list)
        te, _, _ = self._concat(Xv, Yv)
        self.set_data(np.concatenate([tr, te], axis=0), pred_idx, targ_idx)
        n_tr = tr.shape[0]
        return self.fit(np.arange(n_tr), np.arange(n_tr, n_tr + te.shape[0]), epochs, patience, verbose, perm_fn)

    def predict_arrays(self, X_list):
        """``model.predict(X_list)``: list of S arrays [n, O] (a single array when S == 1, multinet.py:279)."""
        mat, pred_idx, targ_idx = self._concat(X_list)
        self.set_data(mat, pred_idx, targ_idx)
        out = self.predict()
        parts = [np.ascontiguousarray(out[:, s * self.O:(s + 1) * self.O]) for s in range(self.S)]
        return parts[0] if self.S == 1 else parts

    # -- persistence (multinet.py:105-124) -------------------------------------------------------------------
    def save(self, path, **extra):
        arrays = {}
        for s, ws in enumerate(self.get_weights()):
            for name, a in zip(("W1", "b1", "W2", "b2"), ws):
                arrays["{}_{}".format(name, s)] = a
        meta = dict(n_pred=np.asarray(self.n_pred), hidden=self.H, sub_outputdim=self.O, batch_size=self.B,
                    learning_rate=self.learning_rate, dropout_rate=self.dropout_rate, seed=self.seed,
                    subnet_ids=np.asarray(self.subnet_ids))
        def portable(a):
            """Labels as an array np.savez can store without pickling, keeping their type: integer gene names stay
            integers (also when they arrive in an object array), everything else becomes text."""
            a = np.asarray(a)
            if a.dtype == object and a.size and all(isinstance(x, (int, np.integer)) and not isinstance(x, bool) for x in a.ravel()):
                return a.astype(np.int64)
            return a.astype(str) if a.dtype.kind in "OUS" else a

        for k, v in extra.items():
            if isinstance(v, (list, tuple)):
                arrays["extra_{}_n".format(k)] = np.asarray(len(v))
                for i, a in enumerate(v):
                    arrays["extra_{}_{}".format(k, i)] = portable(a)
            else:
                arrays["extra_" + k] = portable(v)
        np.savez(path, **arrays, **{"meta_" + k: np.asarray(v) for k, v in meta.items()})

    @staticmethod
    def load(path, math_mode=None, device=None):
        z = np.load(path, allow_pickle=False)
        meta = {k[5:]: z[k] for k in z.files if k.startswith("meta_")}
        eng = Engine([int(p) for p in meta["n_pred"]], hidden=int(meta["hidden"]),
                     sub_outputdim=int(meta["sub_outputdim"]), learning_rate=float(meta["learning_rate"]),
                     batch_size=int(meta["batch_size"]), dropout_rate=float(meta["dropout_rate"]),
                     seed=int(meta["seed"]), math_mode=math_mode, device=device, init_weights=False,
                     subnet_ids=[int(s) for s in meta["subnet_ids"]])
        eng.set_weights([tuple(z["{}_{}".format(n, s)] for n in ("W1", "b1", "W2", "b2")) for s in range(eng.S)])
        extra = {}
        for k in z.files:
            if k.startswith("extra_") and k.endswith("_n"):
                name = k[6:-2]
                extra[name] = [z["extra_{}_{}".format(name, i)] for i in range(int(z[k]))]
        for k in z.files:
            if k.startswith("extra_") and not k.endswith("_n"):
                name = k[6:]
                if name.rsplit("_", 1)[0] not in extra:
                    extra[name] = z[k]
        return eng, extra

    # -- measurement hooks ----------------------------------------------------------------------------------
    def sync(self):
        self._check(self.lib.di_device_sync(self._h))

    def timer_start(self):
        self._check(self.lib.di_timer_start(self._h))

    def timer_stop(self):
        """Device ms since ``timer_start`` measured with CUDA events on the engine's own stream."""
        ms = C.c_float()
        self._check(self.lib.di_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self.lib.di_launch_count(self._h))

    def last_device_ms(self):
        return float(self.lib.di_last_device_ms(self._h))

    def describe(self):
        """Which kernels / knobs the handle runs with (one line) -- goes into the bench line."""
        msg = self.lib.di_describe(self._h)
        return msg.decode() if msg else ""

    def graph_fallbacks(self):
        """Epochs that could not be replayed as a CUDA graph and ran step by step (0 in normal operation)."""
        return int(self.lib.di_graph_fallbacks(self._h))

    def set_profiling(self, on=True):
        self._check(self.lib.di_set_profiling(self._h, 1 if on else 0))

    def kernel_ms(self, which):
        return float(self.lib.di_kernel_ms(self._h, which.encode()))

    def kernel_launches(self, which):
        return int(self.lib.di_kernel_launches(self._h, which.encode()))

    def read_norm(self):
        """The normalised matrix as it sits in HBM, [N, G] float32 (parity check of the device-side log1p)."""
        buf = np.empty((self.n_cells, self.n_genes), np.float32)
        ld = C.c_int64()
        self._check(self.lib.di_debug_read(self._h, b"norm", _lib.f32(buf), buf.size, C.byref(ld)))
        return buf

    def debug_read(self, which):
        ncols = self.S * (self.O if which == "dz2" else self.H) * 2 + 4096
        buf = np.zeros(self.B * ncols, np.float32)
        ld = C.c_int64()
        self._check(self.lib.di_debug_read(self._h, which.encode(), _lib.f32(buf), buf.size, C.byref(ld)))
        return buf[:self.B * ld.value].reshape(self.B, ld.value)
