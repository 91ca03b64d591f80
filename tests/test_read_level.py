"""Read-level consensus network (network half): LatentSpaceLSTM.forward
(medaka/architectures/latent_space_lstm.py:154-207).  Goldens = outputs of the reference's own class on seeded parameters
(tests/golden/make_rl_golden.py).  Bar: probabilities within 2e-5 absolute of the reference (fp32 against fp32 in a different
summation order), labels identical wherever the reference's top-2 margin exceeds 1e-4."""
import os

import numpy as np
import pytest

from oracle import rl_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "rl_forward.npz")
TOL = 2e-5


def _case(g, name):
    seed, B, P, D, dw, gain = g[name + "_args"]
    sd = rl_oracle.synth_rl_state_dict(int(seed), use_dwells=bool(dw), gain=float(gain))
    x = rl_oracle.synth_rl_features(int(B), int(P), int(D), use_dwells=bool(dw), seed=100 + int(seed))
    return sd, x, bool(dw), g[name + "_probs"]


@pytest.mark.parametrize("name", ["small", "deep", "dwells", "hot"])
def test_oracle_matches_reference_class(name):
    g = np.load(GOLD)
    sd, x, dw, want = _case(g, name)
    got = rl_oracle.predict(rl_oracle.build(sd, use_dwells=dw), x)
    assert np.abs(got - want).max() < 2e-6


def _check(got, want):
    assert got.shape == want.shape and np.isfinite(got).all()
    assert np.abs(got - want).max() < TOL, np.abs(got - want).max()
    top2 = np.sort(want, -1)[..., -2:]
    decided = (top2[..., 1] - top2[..., 0]) > 1e-4
    assert np.array_equal(np.argmax(got, -1)[decided], np.argmax(want, -1)[decided])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["small", "deep", "dwells", "hot"])
def test_device_matches_reference_class(name):
    from medaka_b200 import read_level
    g = np.load(GOLD)
    sd, x, dw, want = _case(g, name)
    m = read_level.LatentSpaceLSTM(use_dwells=dw)
    m.load_state_dict(sd)
    _check(m.forward_arrays(x), want)
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("B,P,D", [(1, 17, 1), (9, 65, 5), (3, 1000, 30), (17, 200, 3)])
def test_device_matches_oracle_ragged_shapes(B, P, D):
    """Position counts off the 64-position tile, windows off the 8-window LSTM group, single reads, windows split over
    several device calls."""
    from medaka_b200 import read_level
    sd = rl_oracle.synth_rl_state_dict(5)
    x = rl_oracle.synth_rl_features(B, P, D, seed=B * 1000 + P, empty_rows=min(2, D - 1))
    want = rl_oracle.predict(rl_oracle.build(sd), x)
    m = read_level.LatentSpaceLSTM()
    m.load_state_dict(sd)
    m.max_cells = 40000                      # forces several device calls for the larger shapes
    _check(m.forward_arrays(x), want)
    m.close()


@pytest.mark.gpu
def test_predict_on_batch_interface_and_encoder_check():
    from medaka_b200 import features, read_level, torch_ext
    sd = rl_oracle.synth_rl_state_dict(6)
    m = read_level.LatentSpaceLSTM()
    m.load_state_dict(sd)
    m.check_feature_encoder_compatibility(features.ReadAlignmentFeatureEncoder(include_dwells=False))
    with pytest.raises(ValueError):
        m.check_feature_encoder_compatibility(features.CountsFeatureEncoder())

    class B(object):
        read_level_features = rl_oracle.synth_rl_features(2, 80, 6, seed=3)
    out = m.predict_on_batch(B)
    assert tuple(out.shape) == (2, 80, 5) and abs(float(out.sum(-1).mean()) - 1.0) < 1e-5
    m.close()
