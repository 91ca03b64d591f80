"""Model seam: a drop-in for the reference's GRUModel / TorchModel backed by libmedaka_b200.

Mirrors medaka/architectures/gru.py:13-72 (constructor arguments, ``forward``) and
medaka/models.py:277-383 (``TorchModel``: ``predict_on_batch``, ``device``, ``half``,
``eval``, ``load_state_dict``, ``to_dict``, ``check_feature_encoder_compatibility``),
so ``prediction.run_prediction`` (medaka/prediction.py:44-52) runs unchanged on top of it.
PyTorch tensors are used only as host containers for weights and for the returned CPU
tensor; all arithmetic runs in the CUDA library.  No CPU fallback.
"""
import collections
import inspect
import logging
import warnings

import numpy as np

from medaka_b200 import libmedaka as _lm

ForwardOutput = collections.namedtuple("ForwardOutput", ["probs", "logits", "labels"])

_PRECISIONS = {"tc": 0, "half": 0, "fp32": 1, "full": 1}


def _as_f32(x):
    """numpy float32 C-contiguous view/copy of a torch tensor or array-like."""
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=np.float32)


class PinnedArray(object):
    """A numpy array living in CUDA page-locked host memory (mdk_host_alloc)."""

    def __init__(self, shape, dtype):
        lib = _lm.load()
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        pp = _lm.ffi.new("void **")
        _lm.check(lib.mdk_host_alloc(max(self.nbytes, 1), pp))
        self._ptr = pp[0]
        buf = _lm.ffi.buffer(self._ptr, max(self.nbytes, 1))
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def close(self):
        if getattr(self, "_ptr", None) is not None and _lm.lib is not None:
            self.array = None
            _lm.lib.mdk_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GRUModel(object):
    """Bidirectional GRU consensus model (gru.py:10-72) executing on a B200.

    The constructor signature is the reference's, so ``model_from_dict``-style
    configs (medaka/models.py:392-400) instantiate it directly.
    """

    def __init__(self, num_features=10, num_classes=5, gru_size=128, n_layers=2,
                 bidirectional=True, time_steps=None, classify_activation=None, device=0):
        if time_steps is not None:
            warnings.warn("timesteps is no longer required to be specified")
        if classify_activation is not None:
            warnings.warn("classify_activation is no longer used")
        self.gru_size = gru_size
        self.num_classes = num_classes
        self.num_features = num_features
        self.n_layers = n_layers
        self.bidirectional = bidirectional
        self.normalise = True
        self.half_precision = False
        self.logger = logging.getLogger("B200Model")
        self._device = int(device)
        self._state = {}
        self._pinned = {}
        self._engine = None
        lib = _lm.load()
        _lm.require_gpu(self._device)
        desc = _lm.ffi.new("mdk_model_desc *")
        desc.num_features = num_features
        desc.gru_size = gru_size
        desc.n_layers = n_layers
        desc.bidirectional = 1 if bidirectional else 0
        desc.num_classes = num_classes
        pe = _lm.ffi.new("mdk_engine **")
        _lm.check(lib.mdk_engine_create(self._device, desc, pe))
        self._engine = pe[0]

    # ------------------------------------------------------------------ torch.nn.Module look-alikes
    def eval(self):
        return self

    def to(self, device=None):
        return self

    def device(self):
        """Device where the model has been loaded (medaka/models.py:291-296)."""
        import torch
        return torch.device("cuda", self._device)

    def half(self):
        """TorchModel.half (models.py:298-301): tensor-core path (fp16 hi/lo split operands)."""
        self.half_precision = True
        self.set_precision("tc")
        return self

    def float(self):
        self.half_precision = False
        self.set_precision("fp32")
        return self

    def set_precision(self, mode):
        """'tc' (tcgen05, default) or 'fp32' (CUDA-core FFMA, the --full_precision path)."""
        _lm.check(_lm.lib.mdk_engine_set_precision(self._engine, _PRECISIONS[mode]))

    def parameters(self):
        return iter(self._state.values())

    def count_parameters(self):
        return int(sum(v.size for v in self._state.values()))

    def state_dict(self):
        import torch
        return collections.OrderedDict((k, torch.from_numpy(v.copy())) for k, v in self._state.items())

    def load_state_dict(self, state_dict, strict=True):
        """torch state-dict layout; mirrors datastore.py:150-152."""
        lib = _lm.lib
        H, F = self.gru_size, self.num_features
        ndir = 2 if self.bidirectional else 1
        expected = {}
        for layer in range(self.n_layers):
            n_in = F if layer == 0 else H * ndir
            for sfx in [""] + (["_reverse"] if self.bidirectional else []):
                expected["gru.weight_ih_l%d%s" % (layer, sfx)] = (3 * H, n_in)
                expected["gru.weight_hh_l%d%s" % (layer, sfx)] = (3 * H, H)
                expected["gru.bias_ih_l%d%s" % (layer, sfx)] = (3 * H,)
                expected["gru.bias_hh_l%d%s" % (layer, sfx)] = (3 * H,)
        expected["linear.weight"] = (5, H * ndir)
        expected["linear.bias"] = (5,)
        missing = [k for k in expected if k not in state_dict]
        unexpected = [k for k in state_dict if k not in expected]
        if missing or (strict and unexpected):
            raise RuntimeError("Error(s) in loading state_dict: missing {}, unexpected {}".format(
                missing, unexpected))
        sd = {}
        for k, shape in expected.items():
            a = _as_f32(state_dict[k])
            if a.shape != shape:
                raise RuntimeError("size mismatch for {}: expected {}, got {}".format(k, shape, a.shape))
            sd[k] = a
        ptr = lambda a: _lm.ffi.cast("const float *", _lm.ffi.from_buffer(a))   # noqa: E731
        for layer in range(self.n_layers):
            for d, sfx in enumerate([""] + (["_reverse"] if self.bidirectional else [])):
                _lm.check(lib.mdk_engine_load_gru(
                    self._engine, layer, d,
                    ptr(sd["gru.weight_ih_l%d%s" % (layer, sfx)]), ptr(sd["gru.weight_hh_l%d%s" % (layer, sfx)]),
                    ptr(sd["gru.bias_ih_l%d%s" % (layer, sfx)]), ptr(sd["gru.bias_hh_l%d%s" % (layer, sfx)])))
        _lm.check(lib.mdk_engine_load_linear(self._engine, ptr(sd["linear.weight"]), ptr(sd["linear.bias"])))
        self._state = sd
        return self

    # ------------------------------------------------------------------ TorchModel interface
    def get_model_input_features(self, batch):
        """CountsMatrixModel (medaka/architectures/base_classes.py:9-11)."""
        return batch.counts_matrix

    def check_feature_encoder_compatibility(self, fenc):
        """base_classes.py:13-20: counts-matrix models need a counts encoder of matching width."""
        fvl = getattr(fenc, "feature_vector_length", None)
        if fvl is not None and int(fvl) != int(self.num_features):
            raise ValueError("Feature encoder produces {} features, model expects {}".format(fvl, self.num_features))

    def to_dict(self):
        """models.py:343-361."""
        kwargs = inspect.signature(self.__class__.__init__).parameters
        out = {}
        for k in kwargs:
            if k in ("self", "device"):
                continue
            out[k] = getattr(self, k, kwargs[k].default)
        return {"type": "GRUModel", "kwargs": out}

    def pinned(self, key, shape, dtype):
        """Reusable page-locked staging array, grown on demand."""
        need = int(np.prod(shape))
        cur = self._pinned.get(key)
        if cur is None or cur.dtype != np.dtype(dtype) or int(np.prod(cur.shape)) < need:
            if cur is not None:
                cur.close()
            cur = PinnedArray((need,), dtype)
            self._pinned[key] = cur
        return cur.array[:need].reshape(shape)

    def forward_arrays(self, feats, want_logits=False, want_labels=True):
        """feats float32 [B,T,F] (host) -> ForwardOutput of numpy arrays (copies out of pinned staging)."""
        feats = np.asarray(feats)
        if feats.ndim != 3 or feats.shape[2] != self.num_features:
            raise ValueError("expected features of shape [B, T, {}], got {}".format(self.num_features, feats.shape))
        B, T, F = feats.shape
        self._last_shape = (B, T)
        lib, ffi = _lm.lib, _lm.ffi
        x = self.pinned("feats", (B, T, F), np.float32)
        np.copyto(x, feats, casting="same_kind")
        probs = self.pinned("probs", (B, T, 5), np.float32)
        logits = self.pinned("logits", (B, T, 5), np.float32) if want_logits else None
        labels = self.pinned("labels", (B, T), np.uint8) if want_labels else None
        _lm.check(lib.mdk_engine_forward(
            self._engine, ffi.cast("const float *", ffi.from_buffer(x)), B, T,
            ffi.cast("float *", ffi.from_buffer(probs)),
            ffi.cast("float *", ffi.from_buffer(logits)) if want_logits else ffi.NULL,
            ffi.cast("uint8_t *", ffi.from_buffer(labels)) if want_labels else ffi.NULL))
        return ForwardOutput(probs.copy(), logits.copy() if want_logits else None,
                             labels.copy() if want_labels else None)

    def submit_arrays(self, feats_pinned, probs_out, labels_out=None, logits_out=None):
        """Queue one forward on host arrays WITHOUT waiting (mdk_engine_submit); returns a ticket for ``wait``.

        All arrays must stay alive and untouched until ``wait(ticket)``; page-locked arrays (``pinned``)
        make the copies asynchronous so consecutive calls overlap H2D, compute and D2H.
        """
        B, T, F = feats_pinned.shape
        lib, ffi = _lm.lib, _lm.ffi
        ticket = ffi.new("int64_t *")
        _lm.check(lib.mdk_engine_submit(
            self._engine, ffi.cast("const float *", ffi.from_buffer(feats_pinned)), B, T,
            ffi.cast("float *", ffi.from_buffer(probs_out)),
            ffi.cast("float *", ffi.from_buffer(logits_out)) if logits_out is not None else ffi.NULL,
            ffi.cast("uint8_t *", ffi.from_buffer(labels_out)) if labels_out is not None else ffi.NULL, ticket))
        self._last_shape = (B, T)
        return int(ticket[0])

    def wait(self, ticket):
        _lm.check(_lm.lib.mdk_engine_wait(self._engine, ticket))

    def predict_async(self, batch, slots=2):
        """Asynchronous predict_on_batch: returns a handle whose ``result()`` is the CPU tensor [B,T,5].

        Up to ``slots`` calls may be in flight (each owns one set of page-locked staging arrays).  The engine
        packs consecutive batches of the same window length into one-wave groups window by window (mdk_engine_submit;
        a batch may straddle two groups), so ``run_prediction`` keeps about three groups of batches queued: the
        reference's default 200-window batches then run as 1184-window groups, two computing at a time, with the
        PCIe copies of the neighbouring groups under the compute.
        """
        import torch
        x = _as_f32(self.get_model_input_features(batch))
        B, T, F = x.shape
        slot = getattr(self, "_async_n", 0) % max(int(slots), 1)
        self._async_n = getattr(self, "_async_n", 0) + 1
        xin = self.pinned("afeats%d" % slot, (B, T, F), np.float32)
        np.copyto(xin, x)
        probs = self.pinned("aprobs%d" % slot, (B, T, 5), np.float32)
        labels = self.pinned("alabels%d" % slot, (B, T), np.uint8)
        ticket = self.submit_arrays(xin, probs, labels)
        model = self

        class _Handle(object):
            def result(self_inner):
                model.wait(ticket)
                model.last_labels = labels.copy()
                self_inner.labels = model.last_labels
                return torch.from_numpy(probs.copy())

        return _Handle()

    def lookahead(self, batch_size, window_len=None):
        """How many ``predict_async`` calls of ``batch_size`` windows to keep in flight (four coalesced groups, one per
        big staging lane: two computing, one copying in, one copying out)."""
        pref = self.preferred_batch_size()
        if window_len is not None and batch_size * window_len <= (1 << 18):
            return 14                                   # small forwards rotate over the engine's small lanes
        return int(max(2, min(64, 4 * ((pref + batch_size - 1) // max(batch_size, 1)) + 1)))

    def reserve(self, windows, window_len):
        """Size the engine's compute lanes for coalesced groups of up to ``windows`` windows (mdk_engine_reserve)."""
        _lm.check(_lm.lib.mdk_engine_reserve(self._engine, int(windows), int(window_len)))

    def flush(self):
        _lm.check(_lm.lib.mdk_engine_flush(self._engine))

    def set_products(self, mask):
        """fp16 products per contraction (precision experiments; 7 = fp32-faithful default)."""
        _lm.check(_lm.lib.mdk_engine_set_products(self._engine, int(mask)))

    def set_rec_mode(self, mode):
        """'auto' | 'one' | 'pp': recurrent-kernel selection (mdk_engine_set_rec_mode)."""
        code = {"auto": _lm.lib.MDK_REC_AUTO, "one": _lm.lib.MDK_REC_ONE_TILE, "pp": _lm.lib.MDK_REC_PINGPONG}[mode]
        _lm.check(_lm.lib.mdk_engine_set_rec_mode(self._engine, code))

    def set_group_windows(self, windows):
        _lm.check(_lm.lib.mdk_engine_set_group_windows(self._engine, int(windows)))

    def forward(self, x):
        """gru.py:58-72 on host tensors: returns probabilities (or logits if normalise is off)."""
        import torch
        out = self.forward_arrays(_as_f32(x), want_logits=not self.normalise, want_labels=False)
        return torch.from_numpy(out.probs if self.normalise else out.logits)

    def predict_on_batch(self, batch):
        """TorchModel.predict_on_batch (models.py:303-313): returns a CPU float32 tensor [B,T,5].

        The argmax labels of the same call are kept on ``self.last_labels`` (uint8 [B,T]) so the
        decode stage need not recompute them (north-star: softmax/argmax on the GPU).
        """
        import torch
        out = self.forward_arrays(_as_f32(self.get_model_input_features(batch)), want_logits=False,
                                  want_labels=True)
        self.last_labels = out.labels
        return torch.from_numpy(out.probs)

    # ------------------------------------------------------------------ diagnostics
    def last_timings(self):
        t = _lm.ffi.new("mdk_timings *")
        _lm.check(_lm.lib.mdk_engine_last_timings(self._engine, t))
        return {k: getattr(t, k) for k in ("h2d_ms", "inproj0_ms", "rec0_ms", "inproj1_ms", "rec1_ms",
                                             "head_ms", "d2h_ms", "total_ms", "launches")}

    def read_activation(self, which):
        """Layer output [B,T,256] of the last forward (0 = layer 0, 1 = layer 1) for layer-wise parity."""
        t = self.last_timings()  # syncs
        del t
        shape = self._last_shape
        out = np.empty((shape[0], shape[1], 2 * self.gru_size), dtype=np.float32)
        _lm.check(_lm.lib.mdk_engine_read_activation(
            self._engine, which, _lm.ffi.cast("float *", _lm.ffi.from_buffer(out)), out.size))
        return out

    def launch_count(self):
        return int(_lm.lib.mdk_engine_launch_count(self._engine))

    def keep_activations(self, keep=True):
        """Debugging: keep the layer-1 output in HBM for ``read_activation(1)`` (runs the head as its own kernel)."""
        _lm.check(_lm.lib.mdk_engine_keep_activations(self._engine, 1 if keep else 0))

    def preferred_batch_size(self):
        """Windows per batch that fill the device in one wave (1184 on a B200); ``batch_size="auto"`` in
        ``prediction.run_prediction`` / ``predict_regions`` resolves to this."""
        return int(_lm.lib.mdk_engine_preferred_windows(self._engine))

    @property
    def engine(self):
        return self._engine

    def close(self):
        if self._engine is not None and _lm.lib is not None:
            _lm.lib.mdk_engine_destroy(self._engine)
            self._engine = None
        for p in self._pinned.values():
            p.close()
        self._pinned = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def model_from_dict(d, time_steps=None, device=0):
    """medaka/models.py:392-400 for the architectures this engine implements."""
    name, kwargs = d["type"], dict(d["kwargs"])
    kwargs.pop("read_majority_threshold", None)
    if name == "LatentSpaceLSTM":
        from medaka_b200 import read_level
        return read_level.LatentSpaceLSTM(device=device, **kwargs)
    if name != "GRUModel":
        raise NotImplementedError("medaka_b200 implements GRUModel and LatentSpaceLSTM; got {}".format(name))
    return GRUModel(device=device, **kwargs)


def build_model_torch(feature_len, num_classes, gru_size=128, classify_activation="softmax", time_steps=None,
                      device=0):
    """Legacy model function (medaka/models.py:403-431) -> the same GRUModel."""
    return GRUModel(num_features=feature_len, num_classes=num_classes, gru_size=gru_size, device=device)
