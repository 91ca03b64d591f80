"""Oracle for the model forward: the reference's GRUModel restated in plain torch (CPU fp32).

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows medaka/architectures/gru.py:46-72 (nn.GRU(F,H,2,bidirectional,batch_first)
-> nn.Linear(2H,5) -> softmax(-1)) and medaka/models.py:303-313 (predict_on_batch:
inference_mode, returns a CPU tensor).  On CPU the reference always runs fp32
(medaka/prediction.py:146-148 forces full_precision), which is the parity target.

``manual_forward`` is a second, loop-level restatement of the same arithmetic
(gate order r,z,n; n = tanh(gi_n + r*(gh_n + b_hn))) used to cross-check
intermediate tensors of the CUDA path layer by layer.
"""
import numpy as np
import torch


class GRUOracle(torch.nn.Module):
    """medaka/architectures/gru.py:13-72 without the medaka base classes."""

    def __init__(self, num_features=10, num_classes=5, gru_size=128, n_layers=2,
                 bidirectional=True):
        super().__init__()
        self.gru = torch.nn.GRU(num_features, gru_size, num_layers=n_layers,
                                bidirectional=bidirectional, batch_first=True)
        # gru.py:53-55: the head is hard-coded to 5 outputs
        self.linear = torch.nn.Linear(2 * gru_size if bidirectional else gru_size, 5)
        self.normalise = True

    def forward(self, x, return_logits=False):
        y = self.gru(x)[0]
        logits = self.linear(y)
        probs = torch.softmax(logits, dim=-1)
        if return_logits:
            return probs, logits
        return probs


def build(state_dict, num_features=10, gru_size=128, n_layers=2, bidirectional=True):
    m = GRUOracle(num_features, 5, gru_size, n_layers, bidirectional)
    m.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in state_dict.items()})
    m.eval()
    return m


def predict_on_batch(model, feats, threads=None):
    """feats float32 [B,T,F] (numpy or torch) -> (probs [B,T,5], logits [B,T,5]) numpy fp32.

    Mirrors TorchModel.predict_on_batch (medaka/models.py:303-313) on CPU.
    """
    if threads is not None:
        torch.set_num_threads(int(threads))
    x = torch.as_tensor(feats, dtype=torch.float32)
    with torch.inference_mode():
        probs, logits = model(x, return_logits=True)
    return probs.numpy(), logits.numpy()


def labels_from_probs(probs):
    """argmax as decode_consensus does it (medaka/labels.py:1063): first max wins."""
    return np.argmax(probs, -1).astype(np.uint8)


def manual_forward(state_dict, feats, gru_size=128, n_layers=2):
    """Loop-level fp32 restatement; returns dict of intermediates (numpy).

    out['h0'] / out['h1']: layer outputs [B,T,2H]; out['logits'], out['probs'].
    """
    x = torch.as_tensor(feats, dtype=torch.float32)
    B, T, _ = x.shape
    H = gru_size
    out = {}
    sd = {k: torch.as_tensor(v) for k, v in state_dict.items()}
    inp = x
    for layer in range(n_layers):
        ys = []
        for sfx in ("", "_reverse"):
            w_ih = sd["gru.weight_ih_l%d%s" % (layer, sfx)]
            w_hh = sd["gru.weight_hh_l%d%s" % (layer, sfx)]
            b_ih = sd["gru.bias_ih_l%d%s" % (layer, sfx)]
            b_hh = sd["gru.bias_hh_l%d%s" % (layer, sfx)]
            gi = inp @ w_ih.T + b_ih  # [B,T,3H]
            h = torch.zeros(B, H)
            hs = [None] * T
            order = range(T) if sfx == "" else range(T - 1, -1, -1)
            for t in order:
                gh = h @ w_hh.T + b_hh
                r = torch.sigmoid(gi[:, t, 0:H] + gh[:, 0:H])
                z = torch.sigmoid(gi[:, t, H:2 * H] + gh[:, H:2 * H])
                n = torch.tanh(gi[:, t, 2 * H:] + r * gh[:, 2 * H:])
                h = (1 - z) * n + z * h
                hs[t] = h
            ys.append(torch.stack(hs, 1))
        inp = torch.cat(ys, -1)
        out["h%d" % layer] = inp.numpy()
    logits = inp @ sd["linear.weight"].T + sd["linear.bias"]
    out["logits"] = logits.numpy()
    out["probs"] = torch.softmax(logits, -1).numpy()
    return out
