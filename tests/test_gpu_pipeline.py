"""GPU plumbing test (BASELINE config 1 substitute): the kept orchestration -
region triage -> DataLoader threads -> GPU count normalisation -> Batch.collate -> engine forward
(async look-ahead) -> output store - end to end on synthetic pileups, checked against the oracle."""
import os
import tempfile

import numpy as np
import pytest

from oracle import common_oracle, features_oracle, gru_oracle, synth

pytestmark = pytest.mark.gpu


def _pileup_source(region, bam, encoder):
    # deterministic per region: ~15 % insertion columns, like calculate_pileup would emit
    n_ref = region.end - region.start
    counts, pos = synth.synth_counts(int(n_ref * 1.18), seed=region.start + 17 * len(region.ref_name),
                                     start_major=region.start)
    keep = pos["major"] < region.end
    return [(counts[keep], pos[keep])]


def test_predict_regions_end_to_end():
    from medaka_b200 import common, datastore, features, models, prediction
    sd = synth.synth_state_dict(2)
    model = models.GRUModel(num_features=10)
    model.load_state_dict(sd)
    enc = features.CountsFeatureEncoder(normalise="total", pileup_source=_pileup_source)
    regions = [common.Region("ctgA", 0, 6000), common.Region("ctgB", 100, 3100), common.Region("tiny", 0, 400)]
    chunk_len, ovlp = 1000, 100
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "probs.npzstore")
        prediction.predict_regions(out, None, regions, model, enc, chunk_len=chunk_len, chunk_ovlp=ovlp,
                                   batch_size=4, bam_chunk=100000, bam_workers=2)
        oracle_model = gru_oracle.build(sd)
        expected_names = set()
        with datastore.DataStore(out, "r") as ds:
            for region in regions:
                counts, pos = _pileup_source(region, None, enc)[0]
                feats, depth = features_oracle.post_process_pileup(counts.copy(), pos, "total")
                if len(pos) < chunk_len:
                    ranges = [(0, len(pos))]                      # remainder pass: un-chunked, batch 1
                else:
                    ranges = common_oracle.chunk_ranges(len(pos), chunk_len, ovlp)
                for a, b in ranges:
                    p = pos[a:b]
                    name = "{}:{}.{}-{}.{}".format(region.ref_name, p["major"][0], p["minor"][0],
                                                   p["major"][-1], p["minor"][-1])
                    expected_names.add(name)
                    s = ds.load_sample(name)
                    assert np.array_equal(s.positions, p)
                    assert np.array_equal(np.asarray(s.depth), depth[a:b].astype(np.int64))
                    ref_probs, _ = gru_oracle.predict_on_batch(oracle_model, feats[None, a:b])
                    assert s.label_probs.shape == (b - a, 5) and s.label_probs.dtype == np.float32
                    assert np.abs(s.label_probs - ref_probs[0]).max() < 1e-4
            assert ds.sample_registry == expected_names
    model.close()


def test_submit_wait_matches_synchronous_forward():
    from medaka_b200 import models
    sd = synth.synth_state_dict(4)
    m = models.GRUModel(num_features=10)
    m.load_state_dict(sd)
    batches = [synth.synth_features(5 + i, 200, 10, seed=i) for i in range(5)]
    sync = [m.forward_arrays(b, want_logits=False) for b in batches]
    ins, probs, labels, tickets = [], [], [], []
    for i, b in enumerate(batches):      # each call gets its own pinned buffers so results can be compared at the end
        x = m.pinned("tin%d" % i, b.shape, np.float32)
        np.copyto(x, b)
        p = m.pinned("tp%d" % i, b.shape[:2] + (5,), np.float32)
        l = m.pinned("tl%d" % i, b.shape[:2], np.uint8)
        tickets.append(m.submit_arrays(x, p, l))
        ins.append(x); probs.append(p); labels.append(l)
    for t in tickets:
        m.wait(t)
    for s, p, l in zip(sync, probs, labels):
        assert np.array_equal(s.probs, p) and np.array_equal(s.labels, l)
    m.close()


@pytest.mark.parametrize("group_windows,rec_mode", [(48, "one"), (48, "pp"), (7, "one")])
def test_batches_split_across_groups(group_windows, rec_mode):
    """Batches are packed into groups window by window: with groups smaller than a batch every batch straddles several
    groups (and several lanes); results must equal the one-forward-per-batch results bit for bit and tickets must
    cover the last piece."""
    from medaka_b200 import models
    sd = synth.synth_state_dict(5)
    m = models.GRUModel(num_features=10)
    m.load_state_dict(sd)
    m.set_rec_mode(rec_mode)
    sizes = [70, 33, 1, 120, 16, 50]
    T = 2000                                   # big-lane class: B*T > 2^18 for the larger batches
    batches = [synth.synth_features_fast(n, T, 10, seed=20 + i) for i, n in enumerate(sizes)]
    sync = [m.forward_arrays(b, want_logits=True) for b in batches]
    m.reserve(max(group_windows, 16), T)
    m.set_group_windows(group_windows)
    outs, tickets = [], []
    for i, b in enumerate(batches):
        x = m.pinned("sin%d" % i, b.shape, np.float32)
        np.copyto(x, b)
        p = m.pinned("sp%d" % i, b.shape[:2] + (5,), np.float32)
        l = m.pinned("sl%d" % i, b.shape[:2], np.uint8)
        g = m.pinned("sg%d" % i, b.shape[:2] + (5,), np.float32)
        p[...] = -1.0
        tickets.append(m.submit_arrays(x, p, l, g))
        outs.append((p, l, g))
    for t in reversed(tickets):                # waiting out of order is allowed
        m.wait(t)
    for s, (p, l, g) in zip(sync, outs):
        assert np.array_equal(s.probs, p) and np.array_equal(s.labels, l) and np.array_equal(s.logits, g)
    m.close()
