"""Oracle for the read-level consensus network: LatentSpaceLSTM restated with plain torch modules.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows medaka/architectures/latent_space_lstm.py:34-207 (bidirectional
variant), read_level_modules.py:7-100 (make_1dconv_layers, ReadLevelConv, MeanPooler).  Parameter names are the
reference's state-dict keys, so one state dict drives the reference class (tests/golden/make_rl_golden.py), this
restatement and the engine.
"""
import numpy as np
import torch


def synth_rl_state_dict(seed=0, lstm_size=128, cnn_size=128, use_dwells=False, gain=1.0):
    """Seeded random parameters in torch's own initialisation ranges, with non-trivial BatchNorm statistics."""
    rs = np.random.RandomState(seed)

    def u(shape, bound):
        return torch.from_numpy(rs.uniform(-bound, bound, size=shape).astype(np.float32))
    nin = 6 + 1 + (1 if use_dwells else 0)
    sd = {"base_embedder.weight": torch.from_numpy(rs.randn(6, 6).astype(np.float32)),
          "strand_embedder.weight": torch.from_numpy(rs.randn(3, 6).astype(np.float32))}
    for idx, (cin, k) in (("0", (nin, 1)), ("3", (cnn_size, 17))):
        bound = 1.0 / np.sqrt(cin * k)
        sd["read_level_conv.convs.%s.weight" % idx] = u((cnn_size, cin, k), bound * gain)
        sd["read_level_conv.convs.%s.bias" % idx] = u((cnn_size,), bound)
    for idx in ("2", "5"):
        sd["read_level_conv.convs.%s.weight" % idx] = torch.from_numpy(rs.uniform(0.5, 1.5, cnn_size).astype(np.float32))
        sd["read_level_conv.convs.%s.bias" % idx] = u((cnn_size,), 0.3)
        sd["read_level_conv.convs.%s.running_mean" % idx] = u((cnn_size,), 0.3)
        sd["read_level_conv.convs.%s.running_var" % idx] = torch.from_numpy(rs.uniform(0.3, 1.5, cnn_size).astype(np.float32))
        sd["read_level_conv.convs.%s.num_batches_tracked" % idx] = torch.tensor(7, dtype=torch.long)
    b = 1.0 / np.sqrt(cnn_size)
    sd["read_level_conv.expansion_layer.weight"] = u((lstm_size, cnn_size), b)      # present in the class, unused by forward
    sd["read_level_conv.expansion_layer.bias"] = u((lstm_size,), b)
    sd["pre_pool_expansion_layer.weight"] = u((lstm_size, cnn_size), b)
    sd["pre_pool_expansion_layer.bias"] = u((lstm_size,), b)
    k = 1.0 / np.sqrt(lstm_size)
    for layer in (0, 1):
        cin = lstm_size if layer == 0 else 2 * lstm_size
        for sfx in ("", "_reverse"):
            sd["lstm.weight_ih_l%d%s" % (layer, sfx)] = u((4 * lstm_size, cin), k)
            sd["lstm.weight_hh_l%d%s" % (layer, sfx)] = u((4 * lstm_size, lstm_size), k * gain)
            sd["lstm.bias_ih_l%d%s" % (layer, sfx)] = u((4 * lstm_size,), k)
            sd["lstm.bias_hh_l%d%s" % (layer, sfx)] = u((4 * lstm_size,), k)
    sd["linear.weight"] = u((5, 2 * lstm_size), 1.0 / np.sqrt(2 * lstm_size) * 24)
    sd["linear.bias"] = u((5,), 0.1)
    return sd


def synth_rl_features(B, P, D, use_dwells=False, seed=0, empty_rows=3, ragged=True):
    """Read-level feature tensors int8 [B, P, D, F] the way the featuriser + Batch.collate padding produce them:
    base 0 = no read, 1-4 ACGT, 5 deletion; quality 0-60; strand 0 / 1 (after the clip); mapQ; dwell; trailing reads
    empty (padding to the batch's maximum depth)."""
    rs = np.random.RandomState(seed)
    F = 5 if use_dwells else 4
    x = np.zeros((B, P, D, F), dtype=np.int8)
    for b in range(B):
        depth = D - (rs.randint(0, empty_rows + 1) if empty_rows else 0)
        for d in range(depth):
            lo = rs.randint(0, max(1, P // 3)) if ragged else 0
            hi = P - (rs.randint(0, max(1, P // 3)) if ragged else 0)
            n = hi - lo
            base = rs.choice([1, 2, 3, 4, 5], size=n, p=[0.23, 0.23, 0.23, 0.23, 0.08])
            x[b, lo:hi, d, 0] = base
            x[b, lo:hi, d, 1] = np.where(base == 5, 0, rs.randint(1, 55, n))
            x[b, lo:hi, d, 2] = rs.randint(0, 2)
            x[b, lo:hi, d, 3] = rs.randint(1, 61)
            if use_dwells:
                x[b, lo:hi, d, 4] = np.where(base == 5, 0, rs.randint(1, 30, n))
    return x


class LatentSpaceLSTM(torch.nn.Module):
    def __init__(self, lstm_size=128, cnn_size=128, use_dwells=False, num_classes=5):
        super().__init__()
        self.use_dwells = use_dwells
        self.lstm_size = lstm_size
        nin = 6 + 1 + (1 if use_dwells else 0)
        self.base_embedder = torch.nn.Embedding(6, 6)
        self.strand_embedder = torch.nn.Embedding(3, 6)
        rlc = torch.nn.Module()
        rlc.convs = torch.nn.Sequential(
            torch.nn.Conv1d(nin, cnn_size, kernel_size=1, padding=0), torch.nn.ReLU(), torch.nn.BatchNorm1d(cnn_size),
            torch.nn.Conv1d(cnn_size, cnn_size, kernel_size=17, padding=8), torch.nn.ReLU(), torch.nn.BatchNorm1d(cnn_size))
        rlc.expansion_layer = torch.nn.Linear(cnn_size, lstm_size)
        self.read_level_conv = rlc
        self.pre_pool_expansion_layer = torch.nn.Linear(cnn_size, lstm_size)
        self.lstm = torch.nn.LSTM(lstm_size, lstm_size, num_layers=2, bidirectional=True, batch_first=True)
        self.linear = torch.nn.Linear(2 * lstm_size, num_classes)

    def forward(self, x):
        mask = x.sum((1, -1)) != 0                                       # latent_space_lstm.py:163-165
        e = self.base_embedder(x[:, :, :, 0].long()) + self.strand_embedder(x[:, :, :, 2].long() + 1)
        q = (x[:, :, :, 1] / 25 - 1).unsqueeze(-1)
        parts = [e, q]
        if self.use_dwells:
            parts.append(x[:, :, :, 4].unsqueeze(-1))
        h = torch.cat(parts, dim=-1).permute(0, 2, 3, 1)                 # b, d, f, p
        b, d, f, p = h.shape
        h = self.read_level_conv.convs(h.flatten(0, 1)).permute(0, 2, 1)
        h = self.pre_pool_expansion_layer(h).view(b, d, p, self.lstm_size)
        h = (h * mask[..., None, None]).sum(dim=1) / mask.sum(-1)[..., None, None]      # MeanPooler
        h = self.lstm(h)[0]
        return torch.softmax(self.linear(h), dim=-1)


def build(state_dict, use_dwells=False):
    m = LatentSpaceLSTM(use_dwells=use_dwells)
    m.load_state_dict(state_dict)
    m.eval()
    return m


def predict(model, x, threads=8):
    torch.set_num_threads(threads)
    with torch.inference_mode():
        return model(torch.from_numpy(np.asarray(x))).numpy()
