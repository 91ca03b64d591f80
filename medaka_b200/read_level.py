"""Read-level consensus model behind the reference's model seam.

Mirror of ``medaka.architectures.latent_space_lstm.LatentSpaceLSTM`` (latent_space_lstm.py:34-207) as far as inference
needs it: constructor keywords, ``load_state_dict`` with the reference's parameter names, ``predict_on_batch`` on a batch
whose ``read_level_features`` is the int8 tensor [batch, positions, reads, features] of
``ReadAlignmentFeatureEncoder`` (``ReadLevelFeaturesModel.get_model_input_features``, base_classes.py:29-36), the
compatibility check of :209-236.  The arithmetic runs in libmedaka_b200 (``mdk_rl_*``, csrc/readlevel.cu).
"""
import warnings

import numpy as np

from medaka_b200 import libmedaka as _lm


class LatentSpaceLSTM(object):

    def __init__(self, num_classes=5, lstm_size=128, cnn_size=128, kernel_sizes=(1, 17), pooler_type="mean",
                 pooler_args=None, use_dwells=False, bases_alphabet_size=6, bases_embedding_size=6, bidirectional=True,
                 time_steps=None, device=0):
        if time_steps is not None:
            warnings.warn("timesteps is no longer required to be specified")
        if list(kernel_sizes) != [1, 17] or pooler_type != "mean" or not bidirectional or bases_alphabet_size != 6 \
                or bases_embedding_size != 6:
            raise NotImplementedError("the engine implements the bidirectional, mean-pooled, kernel_sizes=[1, 17] model")
        self.num_classes, self.lstm_size, self.cnn_size = num_classes, lstm_size, cnn_size
        self.kernel_sizes, self.pooler_type, self.pooler_args = list(kernel_sizes), pooler_type, pooler_args or {}
        self.use_dwells = use_dwells
        self.bases_alphabet_size, self.bases_embedding_size = bases_alphabet_size, bases_embedding_size
        self.bidirectional = bidirectional
        self.normalise = True
        self.device_index = device
        lib, ffi = _lm.load(), _lm.ffi
        _lm.require_gpu(device)
        pe = ffi.new("mdk_rl_engine **")
        _lm.check(lib.mdk_rl_create(device, lstm_size, cnn_size, 1 if use_dwells else 0, num_classes, pe))
        self._engine = pe[0]
        self.max_cells = 1 << 26            # positions x reads per device call (bounds the per-call device scratch)

    # ---- TorchModel interface (medaka/models.py:233-313) ----
    def load_state_dict(self, state_dict, strict=True):
        lib, ffi = _lm.lib, _lm.ffi
        for name, t in state_dict.items():
            if name.endswith("num_batches_tracked"):
                continue
            a = np.ascontiguousarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t, dtype=np.float32)
            _lm.check(lib.mdk_rl_load(self._engine, name.encode(), ffi.cast("const float *", ffi.from_buffer(a)), a.size))
        return self

    def set_conv(self, tensor_cores=True, lstm_tensor_cores=None):
        """k = 17 convolution / LSTM recurrences on the tensor cores (default) or on the fp32 CUDA cores (validation)."""
        if lstm_tensor_cores is None:
            lstm_tensor_cores = tensor_cores
        _lm.check(_lm.lib.mdk_rl_set_conv(self._engine, (1 if tensor_cores else 0) | (2 if lstm_tensor_cores else 0)))

    def eval(self):
        return self

    def half(self):
        return self

    def device(self):
        return "cuda:%d" % self.device_index

    def get_model_input_features(self, batch):
        return batch.read_level_features

    def check_feature_encoder_compatibility(self, fenc):
        """latent_space_lstm.py:209-236."""
        from medaka_b200 import features
        if not isinstance(fenc, features.ReadAlignmentFeatureEncoder):
            raise ValueError("LatentSpaceLSTM expects a ReadAlignmentFeatureEncoder.")
        if len(fenc.dtypes) > 1:
            raise NotImplementedError("LatentSpaceLSTM is currently only implemented for one dtype.")
        if self.use_dwells and not getattr(fenc, "include_dwells", False):
            raise ValueError("Model expects dwells, however include_dwells not set in the feature encoder.")

    def forward_arrays(self, x):
        """x int8 [B, P, D, F] -> probabilities float32 [B, P, 5]."""
        x = np.ascontiguousarray(x.detach().cpu().numpy() if hasattr(x, "detach") else x)
        if x.dtype != np.int8:
            x = x.astype(np.int8)
        if x.ndim != 4:
            raise ValueError("expected read-level features [batch, positions, reads, features]")
        B, P, D, F = x.shape
        lib, ffi = _lm.lib, _lm.ffi
        probs = np.empty((B, P, 5), dtype=np.float32)
        step = max(1, int(self.max_cells // max(P * D, 1)))
        for b0 in range(0, B, step):
            xb = np.ascontiguousarray(x[b0:b0 + step])
            pb = probs[b0:b0 + step]
            _lm.check(lib.mdk_rl_forward(self._engine, ffi.cast("const int8_t *", ffi.from_buffer(xb)), len(xb), P, D, F,
                                         ffi.cast("float *", ffi.from_buffer(pb))))
        return probs

    def predict_on_batch(self, batch):
        import torch
        return torch.from_numpy(self.forward_arrays(self.get_model_input_features(batch)))

    def to_dict(self):
        return {"type": "LatentSpaceLSTM", "kwargs": {
            "num_classes": self.num_classes, "lstm_size": self.lstm_size, "cnn_size": self.cnn_size,
            "kernel_sizes": self.kernel_sizes, "pooler_type": self.pooler_type, "pooler_args": self.pooler_args,
            "use_dwells": self.use_dwells, "bases_alphabet_size": self.bases_alphabet_size,
            "bases_embedding_size": self.bases_embedding_size, "bidirectional": self.bidirectional}}

    def close(self):
        if getattr(self, "_engine", None) is not None and _lm.lib is not None:
            _lm.lib.mdk_rl_destroy(self._engine)
            self._engine = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
