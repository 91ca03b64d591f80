"""Pin the CPU oracle (oracle/) to the reference: committed outputs of the REAL reference
classes (tests/golden/make_golden.py) and the reference's own literal golden vectors."""
import os

import numpy as np
import pytest

from oracle import common_oracle, features_oracle, gru_oracle, labels_oracle, pileup_oracle, synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


@pytest.mark.parametrize("case", ["small", "long", "hot", "f20", "b1"])
def test_gru_oracle_matches_reference_outputs(golden_dir, case):
    g = _load(golden_dir, "gru_forward.npz")
    seed, B, T, F, head_gain, rec_gain = g[case + "_args"]
    sd = synth.synth_state_dict(int(seed), num_features=int(F), head_gain=head_gain, rec_gain=rec_gain)
    feats = synth.synth_features(int(B), int(T), int(F), seed=100 + int(seed))
    model = gru_oracle.build(sd, num_features=int(F))
    probs, logits = gru_oracle.predict_on_batch(model, feats)
    # same torch kernels as the reference's own classes -> tight tolerance (thread-count reordering only)
    np.testing.assert_allclose(probs, g[case + "_probs"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(logits, g[case + "_logits"], rtol=0, atol=2e-5)
    assert np.array_equal(np.argmax(probs, -1), np.argmax(g[case + "_probs"], -1))


def test_manual_forward_matches_nn_gru():
    sd = synth.synth_state_dict(11)
    feats = synth.synth_features(2, 40, 10, seed=3)
    model = gru_oracle.build(sd)
    probs, logits = gru_oracle.predict_on_batch(model, feats)
    man = gru_oracle.manual_forward(sd, feats)
    np.testing.assert_allclose(man["logits"], logits, atol=2e-5, rtol=0)
    np.testing.assert_allclose(man["probs"], probs, atol=2e-6, rtol=0)


def test_norm_indices(golden_dir):
    g = _load(golden_dir, "norm_indices.npz")
    # medaka/test/test_counts.py:486-513 literals
    assert features_oracle.pileup_counts_norm_indices([""]) == {
        ("", True): [0, 1, 2, 3, 8], ("", False): [4, 5, 6, 7, 9]}
    assert features_oracle.pileup_counts_norm_indices(["1", "2"]) == {
        ("1", True): [0, 1, 2, 3, 8], ("2", True): [10, 11, 12, 13, 18],
        ("1", False): [4, 5, 6, 7, 9], ("2", False): [14, 15, 16, 17, 19]}
    for key in g.files:
        if key == "meta":
            continue
        dtypes, nq, dt, rev = key.split("|")
        got = features_oracle.pileup_counts_norm_indices(dtypes.split(","), int(nq))
        assert got[dt, bool(int(rev))] == list(g[key])


NORM_CASES = ["simple", "synth", "synth_minor_start", "synth2dt", "deep"]


@pytest.mark.parametrize("name", NORM_CASES)
@pytest.mark.parametrize("norm", ["total", "fwd_rev", None])
@pytest.mark.parametrize("sym", [False, True])
def test_post_process_matches_reference(golden_dir, name, norm, sym):
    g = _load(golden_dir, "post_process.npz")
    counts = g[name + "_counts"].copy()
    pos = np.empty(len(counts), dtype=[("major", "<i8"), ("minor", "<i8")])
    pos["major"], pos["minor"] = g[name + "_major"], g[name + "_minor"]
    dtypes = ("r9", "r10") if name == "synth2dt" else ("",)
    feats, depth = features_oracle.post_process_pileup(counts, pos, norm, dtypes, sym)
    key = "%s_%s_%d" % (name, norm, int(sym))
    assert feats.dtype == np.float32
    assert np.array_equal(feats, g[key + "_features"])
    assert np.array_equal(depth, g[key + "_depth"])


def test_post_process_reference_literal_goldens():
    """medaka/test/test_counts.py:92-115 ('total') and :152-174 (sym_indels)."""
    g_counts = np.array(
        [[2, 0, 0, 0, 2, 0, 0, 0, 0, 0], [0, 2, 0, 0, 0, 2, 0, 0, 0, 0],
         [2, 0, 0, 0, 2, 0, 0, 0, 0, 0], [0, 1, 0, 1, 0, 0, 0, 1, 0, 1],
         [1, 0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 2, 0, 0, 0, 2, 0, 0, 0],
         [2, 0, 0, 0, 2, 0, 0, 0, 0, 0], [0, 0, 0, 2, 0, 0, 0, 2, 0, 0],
         [0, 0, 2, 0, 0, 0, 2, 0, 0, 0]], dtype=np.uint64)
    pos = np.array([(0, 0), (1, 0), (2, 0), (3, 0), (3, 1), (4, 0), (5, 0), (6, 0), (7, 0)],
                   dtype=[("major", "<i8"), ("minor", "<i8")])
    expected = np.array([
        [0.5, 0., 0., 0., 0.5, 0., 0., 0., 0., 0.], [0., 0.5, 0., 0., 0., 0.5, 0., 0., 0., 0.],
        [0.5, 0., 0., 0., 0.5, 0., 0., 0., 0., 0.], [0., 0.25, 0., 0.25, 0., 0., 0., 0.25, 0., 0.25],
        [0.25, 0., 0., 0., 0., 0., 0., 0., 0., 0.], [0., 0., 0.5, 0., 0., 0., 0.5, 0., 0., 0.],
        [0.5, 0., 0., 0., 0.5, 0., 0., 0., 0., 0.], [0., 0., 0., 0.5, 0., 0., 0., 0.5, 0., 0.],
        [0., 0., 0.5, 0., 0., 0., 0.5, 0., 0., 0.]], dtype=np.float32)
    feats, _ = features_oracle.post_process_pileup(g_counts.copy(), pos, "total")
    assert np.array_equal(feats, expected)
    expected[4] = [0.25, 0., 0., 0., 0., 0., 0., 0., 0.25, 0.5]
    feats, _ = features_oracle.post_process_pileup(g_counts.copy(), pos, "total", sym_indels=True)
    assert np.array_equal(feats, expected)


def test_chunks_and_region_split(golden_dir):
    g = _load(golden_dir, "chunks.npz")
    for key in g.files:
        parts = key.split("_")
        if parts[0] == "chunks":
            n, cl, ov = map(int, parts[1:])
            got = np.array(common_oracle.chunk_ranges(n, cl, ov), dtype=np.int64).reshape(-1, 2)
            assert np.array_equal(got, g[key]), key
        elif parts[0] == "split":
            start, end, size, ov, fixed = map(int, parts[1:])
            got = np.array(common_oracle.region_split(start, end, size, ov, bool(fixed)), dtype=np.int64)
            assert np.array_equal(got, g[key]), key
    assert [len(x) for x in common_oracle.grouper(range(10), 4)] == list(g["grouper_10_4"])


def test_decode_consensus(golden_dir):
    # medaka/test/test_labels.py:239-266
    p = np.array([[0, 1, 0, 0, 0], [0, 0, 1, 0, 0], [0, 0, 1, 0, 0], [1, 0, 0, 0, 0],
                  [0, 0, 0, 0, 1], [0, 0, 0, 1, 0], [0, 0, 0, 1, 0], [1, 0, 0, 0, 0]])
    assert labels_oracle.decode_consensus(p) == "ACCTGG"
    p = np.array([[0., 0.991, 0.009, 0., 0.], [0.1, 0., 0.9, 0., 0.], [0.9, 0., 0.02, 0.04, 0.04],
                  [0, 0, 0, 0, 1], [0, 0.1, 0.1, 0.6, 0.2], [0, 0.01, 0.1, 0.88, 0.01]])
    assert labels_oracle.decode_consensus(p, with_qualities=True) == ("ACTGG", "5+g$*")
    g = _load(golden_dir, "decode.npz")
    seq, qual = labels_oracle.decode_consensus(g["probs"], with_qualities=True)
    assert seq.encode() == g["seq"].tobytes() and qual.encode() == g["qual"].tobytes()
    seq, qual = labels_oracle.decode_consensus(g["probs"], with_gaps=True, with_qualities=True)
    assert seq.encode() == g["seq_gaps"].tobytes() and qual.encode() == g["qual_gaps"].tobytes()
    labels, q = labels_oracle.decode_arrays(g["probs"])
    assert bytes(np.array([ord(c) for c in "*ACGT"], dtype="u1")[labels]) == g["seq_gaps"].tobytes()
    assert q.tobytes() == g["qual_gaps"].tobytes()


# ------------------------------------------------------------------ pileup counts
# medaka/test/mock_data.py:22-100 (reads) and medaka/test/test_counts.py:298-311 (expected)
SIMPLE_CALLS = [
    dict(query_name="basecall_1", seq="ACATGATG", cigar="8=", mapq=40, flag=0, pos=0,
         tags={"AA": 1, "DT": "r9"}),
    dict(query_name="basecall_2", seq="ACAGATG", cigar="3=1D4=", mapq=10, flag=0, pos=0,
         tags={"AA": 1, "DT": "r9"}),
    dict(query_name="basecall_3", seq="ACATAGATG", cigar="4=1I4=", mapq=16, flag=16, pos=0,
         tags={"AA": 2, "DT": "r9"}),
    dict(query_name="basecall_4", seq="ACACGATG", cigar="3=1X4=", mapq=24, flag=16, pos=0,
         tags={"DT": "r10"}),
]
EXPECTED_COUNTS = np.array(
    [[2, 0, 0, 0, 2, 0, 0, 0, 0, 0], [0, 2, 0, 0, 0, 2, 0, 0, 0, 0], [2, 0, 0, 0, 2, 0, 0, 0, 0, 0],
     [0, 1, 0, 1, 0, 0, 0, 1, 0, 1], [1, 0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 2, 0, 0, 0, 2, 0, 0, 0],
     [2, 0, 0, 0, 2, 0, 0, 0, 0, 0], [0, 0, 0, 2, 0, 0, 0, 2, 0, 0], [0, 0, 2, 0, 0, 0, 2, 0, 0, 0]],
    dtype=np.uint64)
EXPECTED_POS = np.array([(0, 0), (1, 0), (2, 0), (3, 0), (3, 1), (4, 0), (5, 0), (6, 0), (7, 0)],
                        dtype=[("major", "<i8"), ("minor", "<i8")])


def test_pileup_counts_simple_golden():
    counts, pos = pileup_oracle.pileup_counts(SIMPLE_CALLS, 0, 8)
    assert np.array_equal(counts, EXPECTED_COUNTS)
    assert np.array_equal(pos, EXPECTED_POS)


def test_pileup_counts_tag_filters():
    # test_counts.py:320-334: AA=1 -> 2 reads; keep_missing -> 3 reads
    c, _ = pileup_oracle.pileup_counts(SIMPLE_CALLS, 0, 8, tag_name="AA", tag_value=1)
    assert set(c.sum(axis=1)) == {2}
    c, _ = pileup_oracle.pileup_counts(SIMPLE_CALLS, 0, 8, tag_name="AA", tag_value=1, keep_missing=True)
    assert set(c.sum(axis=1)) == {3}


def test_pileup_counts_dtypes_and_read_groups():
    c, _ = pileup_oracle.pileup_counts(SIMPLE_CALLS, 0, 8, dtypes=["r9", "r10"])
    assert c.shape == (9, 20)                                  # test_counts.py:348-353
    reads = []
    for rg in ("first", "second"):
        for r in SIMPLE_CALLS:
            r2 = dict(r, tags=dict(r["tags"], RG=rg), query_name=r["query_name"] + "_" + rg)
            reads.append(r2)
    c, p = pileup_oracle.pileup_counts(reads, 0, 8)            # test_counts.py:355-380
    assert np.array_equal(c, 2 * EXPECTED_COUNTS) and np.array_equal(p, EXPECTED_POS)
    for rg in ("first", "second"):
        c, p = pileup_oracle.pileup_counts(reads, 0, 8, read_group=rg)
        assert np.array_equal(c, EXPECTED_COUNTS)
    c, p = pileup_oracle.pileup_counts(reads, 0, 8, read_group="nonsense")
    assert len(c) == 0


def test_pileup_counts_chunked_equals_unchunked():
    # test_counts.py:383-392: region_split=3 then contiguity enforcement gives the same answer
    parts = [pileup_oracle.pileup_counts(SIMPLE_CALLS, s, e) for s, e in
             common_oracle.region_split(0, 8, 3, fixed_size=False)]
    chunks = features_oracle.enforce_pileup_chunk_contiguity(parts)
    assert len(chunks) == 1
    assert np.array_equal(chunks[0][0], EXPECTED_COUNTS) and np.array_equal(chunks[0][1], EXPECTED_POS)


def test_contiguity_splits_on_gaps():
    # test_counts.py:229-243 shape of behaviour: a coverage gap splits the chunk
    reads = [dict(SIMPLE_CALLS[0], pos=0), dict(SIMPLE_CALLS[0], pos=20)]
    c, p = pileup_oracle.pileup_counts(reads, 0, 100)
    chunks = features_oracle.enforce_pileup_chunk_contiguity([(c, p)])
    assert [len(x[1]) for x in chunks] == [8, 8]


# medaka/test/test_labels.py:101-135 (VariantBoundaries.test_001_boundaries)
VARIANT_CASES = [
    ([0, 0, 0, 0], 'ATCG', 'ATCG', '----'), ([0, 1, 0, 0], 'A*CG', 'ATCG', '-+--'),
    ([0, 1, 2, 0], 'A**G', 'ATCG', '-++-'), ([0, 1, 2, 3], 'A***', 'ATC*', '-+++'),
    ([0, 1, 2, 3], 'A***', 'A*CG', '-+++'), ([0, 0, 0, 0], 'ATCG', 'AACG', '-+--'),
    ([0, 0, 0, 0], 'ATCG', 'ATCC', '---+'), ([0, 1, 2, 3], 'A***', 'CTC*', '++++'),
    ([0, 1, 2, 3], 'A***', 'C*CG', '++++'), ([0, 1, 2, 3, 0], 'A***A', 'CTC*A', '++++-'),
    ([0, 1, 2, 3, 0], 'A***A', 'C*CGG', '+++++'), ([0, 1, 2, 3, 0, 1, 2], 'A***A**', 'CTC*A*A', '++++-++'),
    ([0, 1, 2, 3, 0, 1, 2], 'A***A**', 'C*CGGA*', '+++++++'),
]


def test_variant_columns_reference_cases():
    for minor, ref, pred, exp in VARIANT_CASES:
        got = labels_oracle.variant_columns(np.array(minor), np.array(list(ref)), np.array(list(pred)))
        assert "".join("+" if x else "-" for x in got) == exp, (minor, ref, pred)
