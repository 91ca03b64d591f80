"""Golden outputs of the REFERENCE's read-level network (build container only).

Run:  python tests/golden/make_rl_golden.py     (needs /root/reference; writes tests/golden/rl_forward.npz)

Imports medaka.architectures.latent_space_lstm.LatentSpaceLSTM UNMODIFIED (behind the inert stand-ins of make_golden.py
for the absent third-party modules), loads seeded parameters (oracle/rl_oracle.py::synth_rl_state_dict: the state-dict
keys are the reference class's own), puts it in eval mode and records `predict_on_batch`-style outputs for seeded
read-level feature tensors.  The restatement in oracle/rl_oracle.py is asserted against them here.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden  # noqa: E402
from oracle import rl_oracle  # noqa: E402

CASES = {            # name: (seed, B, P, D, use_dwells, gain)
    "small": (0, 2, 96, 7, False, 1.0),
    "deep": (1, 1, 300, 40, False, 1.0),
    "dwells": (2, 3, 130, 9, True, 1.0),
    "hot": (3, 2, 500, 12, False, 2.5),
}


def main():
    make_golden.install_stubs()
    sys.path.insert(0, "/root/reference")
    from medaka.architectures.latent_space_lstm import LatentSpaceLSTM
    out = {}
    for name, (seed, B, P, D, dw, gain) in CASES.items():
        sd = rl_oracle.synth_rl_state_dict(seed, use_dwells=dw, gain=gain)
        ref = LatentSpaceLSTM(use_dwells=dw)
        missing = ref.load_state_dict(sd)
        ref.eval()
        x = rl_oracle.synth_rl_features(B, P, D, use_dwells=dw, seed=100 + seed)
        torch.set_num_threads(8)
        with torch.inference_mode():
            probs = ref(torch.from_numpy(x)).numpy()
        mine = rl_oracle.predict(rl_oracle.build(sd, use_dwells=dw), x)
        err = float(np.abs(mine - probs).max())
        assert err < 2e-6, (name, err)
        out[name + "_args"] = np.array([seed, B, P, D, int(dw), gain], dtype=np.float64)
        out[name + "_probs"] = probs
        print(name, probs.shape, "restatement vs reference %.2e" % err, "mean max prob %.3f" % probs.max(-1).mean())
    np.savez_compressed(os.path.join(HERE, "rl_forward.npz"), **out)


if __name__ == "__main__":
    main()
