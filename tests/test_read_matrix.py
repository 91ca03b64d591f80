"""Read-level featuriser (first function): calculate_read_alignment
(src/medaka_read_matrix.c:277-615) and the chunk join of medaka/features.py:411-557.

CPU: the oracle against the reference's literal expectations (medaka/test/test_read_alignment_matrix.py:103-259, parsed
out of the test source by tests/golden/make_read_matrix_golden.py, which also asserts the real-BAM regression shapes of
:27-71 on the oracle) and against the stored real-BAM slice; the chunk join against the output of the reference's own
function.  GPU: mdk_read_matrix against the oracle, bit for bit.
"""
import os

import numpy as np
import pytest

from oracle import pileup_oracle, read_matrix_oracle, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "read_matrix.npz")

# medaka/test/mock_data.py:32-100 (the four basecalls; quality of the first one dropped as the reference's test does)
MOCK = [
    dict(query_name="basecall_1", seq="ACATGATG", cigar="8=", mapq=40, flag=0, pos=0, qual=None,
         tags={"DT": "r9", "mv": [5, 1, 0, 0, 1, 0, 1, 1, 1, 0, 0, 0, 1, 0, 1, 0, 1, 0, 0]}),
    dict(query_name="basecall_2", seq="ACAGATG", cigar="3=1D4=", mapq=10, flag=0, pos=0, qual=[0, 1, 4, 1, 1, 1, 2],
         tags={"DT": "r9", "mv": [5, 1, 1, 0, 0, 1, 1, 0, 0, 0, 1, 0, 1, 1, 0, 0]}),
    dict(query_name="basecall_3", seq="ACATAGATG", cigar="4=1I4=", mapq=16, flag=16, pos=0,
         qual=[2, 1, 4, 5, 1, 1, 1, 2, 1],
         tags={"DT": "r9", "mv": [5, 1, 0, 1, 0, 1, 0, 1, 1, 0, 0, 0, 1, 0, 1, 1, 0, 0, 1, 0]}),
    dict(query_name="basecall_4", seq="ACACGATG", cigar="3=1X4=", mapq=24, flag=16, pos=0, qual=[2, 1, 4, 1, 1, 1, 2, 1],
         tags={"DT": "r10", "mv": [5, 1, 0, 1, 1, 0, 1, 0, 1, 0, 1, 0, 0, 0, 1, 0, 1, 1, 0, 0, 1, 0, 1, 0]}),
]


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _slice_batch(g):
    from medaka_b200 import bam
    names = str(g["names"]).split("\n")
    return bam.RecordBatch(pos=g["pos"], flag=g["flag"], mapq=g["mapq"], dtype=np.zeros(len(g["pos"]), np.uint8),
                           cigar=g["cigar"], cigar_off=g["cigar_off"], seq=g["seq"], seq_off=g["seq_off"],
                           l_seq=g["l_seq"], names=names, tags=None, qual=g["qual"], aux=g["aux"], aux_off=g["aux_off"])


def _records(batch):
    from medaka_b200 import bam
    recs = pileup_oracle.records_from_batch(batch)
    qoff = np.concatenate([[0], np.cumsum(batch.l_seq)])
    for i, r in enumerate(recs):
        r["query_name"] = batch.names[i]
        r["qual"] = batch.qual[qoff[i]:qoff[i + 1]].tolist()
        r["tags"] = bam._parse_tags(bytes(batch.aux[int(batch.aux_off[i]):int(batch.aux_off[i + 1])]), arrays=True)
    return recs


KINDS = {"default": {}, "rpr": dict(row_per_read=True, max_reads=1000), "max5": dict(max_reads=5),
         "hap": dict(include_haplotype=True)}


def test_oracle_reproduces_reference_literals(gold):
    m, pos, _, _ = read_matrix_oracle.read_alignment(MOCK, 0, 100)
    assert np.array_equal(m, gold["mock_plain"])
    assert pos["major"].tolist() == [0, 1, 2, 3, 3, 4, 5, 6, 7] and pos["minor"].tolist() == [0, 0, 0, 0, 1, 0, 0, 0, 0]
    md, _, _, _ = read_matrix_oracle.read_alignment(MOCK, 0, 100, include_dwells=True)
    assert np.array_equal(md, gold["mock_dwell"])


def test_oracle_matches_stored_real_slice(gold):
    recs = _records(_slice_batch(gold))
    for key, kw in KINDS.items():
        m, pos, left, right = read_matrix_oracle.read_alignment(recs, int(gold["start"]), int(gold["end"]), **kw)
        assert np.array_equal(m, gold["mat_" + key]), key
        assert np.array_equal(pos["major"], gold["major"]) and np.array_equal(pos["minor"], gold["minor"])
        assert [str(x) for x in gold["left_" + key]] == left and [str(x) for x in gold["right_" + key]] == right


def test_chunk_join_matches_reference_function(gold):
    from medaka_b200 import features
    chunks = []
    for k in range(3):
        pos = np.empty(len(gold["chunk%d_major" % k]), dtype=[("major", "<i8"), ("minor", "<i8")])
        pos["major"], pos["minor"] = gold["chunk%d_major" % k], gold["chunk%d_minor" % k]
        chunks.append((gold["chunk%d_mat" % k], pos, (gold["chunk%d_left" % k], gold["chunk%d_right" % k])))
    joined = features._join_read_matrix_chunks(chunks)
    assert len(joined) == 1
    assert np.array_equal(joined[0][0], gold["joined_mat"]) and np.array_equal(joined[0][1]["major"], gold["joined_major"])
    # a coverage gap splits a sub-region result and nothing is joined across it
    m, p, ids = chunks[0]
    keep = np.ones(len(p), dtype=bool)
    keep[40:60] = False
    keep[p["major"] == p["major"][39]] = keep[39]
    parts = features._join_read_matrix_chunks([(m[keep], p[keep], ids)])
    assert len(parts) == 2 and sum(len(x[1]) for x in parts) == int(keep.sum())


def test_chunk_join_hard_case_matches_reference_function(gold):
    """Five sub-regions of unequal depth with reads entering and leaving and a coverage gap inside the third."""
    from medaka_b200 import features
    chunks = []
    for k in range(5):
        pos = np.empty(len(gold["hard%d_major" % k]), dtype=[("major", "<i8"), ("minor", "<i8")])
        pos["major"], pos["minor"] = gold["hard%d_major" % k], gold["hard%d_minor" % k]
        chunks.append((gold["hard%d_mat" % k], pos, (gold["hard%d_left" % k], gold["hard%d_right" % k])))
    joined = features._join_read_matrix_chunks(chunks)
    assert len(joined) == int(gold["hard_n"]) >= 2
    for i, (m, p) in enumerate(joined):
        assert np.array_equal(m, gold["hard_joined%d_mat" % i]), i
        assert np.array_equal(p["major"], gold["hard_joined%d_major" % i])


# ---------------------------------------------------------------------------------------------------- GPU
def _device(batch, start, end, **kw):
    from medaka_b200 import features
    m, pos, (left, right) = features.read_matrix_from_batch(batch, start, end, **kw)
    return np.maximum(m, 0), pos, [x.decode() for x in left], [x.decode() for x in right]


@pytest.mark.gpu
def test_device_matches_reference_literals(gold):
    from medaka_b200 import bam
    batch = bam.records_from_dicts(MOCK)
    m, pos, _, _ = _device(batch, 0, 100)
    assert np.array_equal(m, gold["mock_plain"])
    md, _, _, _ = _device(batch, 0, 100, include_dwells=True)
    assert np.array_equal(md, gold["mock_dwell"])
    mt, _, _, _ = _device(bam.records_from_dicts(MOCK, dtypes=("r9", "r10")), 0, 100, num_dtypes=2, include_dwells=True,
                          include_haplotype=True)
    want, _, _, _ = read_matrix_oracle.read_alignment(MOCK, 0, 100, dtypes=("r9", "r10"), include_dwells=True,
                                                       include_haplotype=True)
    assert mt.shape == (9, 4, 7) and np.array_equal(mt, want)


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(KINDS))
def test_device_matches_real_slice(gold, key):
    batch = _slice_batch(gold)
    m, pos, left, right = _device(batch, int(gold["start"]), int(gold["end"]), **KINDS[key])
    assert np.array_equal(m, gold["mat_" + key])
    assert np.array_equal(pos["major"], gold["major"]) and np.array_equal(pos["minor"], gold["minor"])
    assert left == [str(x) for x in gold["left_" + key]] and right == [str(x) for x in gold["right_" + key]]


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n_reads,kw", [
    (1, 260, {}), (2, 260, dict(row_per_read=True, max_reads=400)), (3, 260, dict(max_reads=7)),
    (4, 260, dict(include_dwells=True, include_haplotype=True)), (5, 260, dict(max_reads=150)),
    (5, 700, dict(max_reads=100)),                          # deeper than the row budget: reads without a row are dropped
    (5, 700, dict(row_per_read=True, max_reads=300)),       # the row buffer grows (medaka_read_matrix.c:359-371) and fills up
])
def test_device_matches_oracle_on_synthetic_reads(seed, n_reads, kw):
    """Random CIGARs with insertions after deletions, consecutive insertions, soft clips, filtered reads, move tables and
    haplotags; region cut inside reads.  (No reference skips: the reference's row bookkeeping measures a read's end
    without them, medaka_read_matrix.c:258-273, so a spliced read's row can be handed to another read while it is still
    listed and the two then share the row's struct - not a behaviour to pin.)"""
    from medaka_b200 import bam
    rs = np.random.RandomState(seed)
    recs = synth.synth_reads(n_reads, 3000, seed=seed, mean_len=500)
    recs.sort(key=lambda r: r["pos"])
    for i, r in enumerate(recs):
        r["query_name"] = "read_%d" % i
        r["qual"] = rs.randint(0, 60, len(r["seq"])).tolist() if rs.uniform() < 0.9 else None
        tags = {}
        if rs.uniform() < 0.7:
            tags["HP"] = int(rs.randint(0, 3))
        if rs.uniform() < 0.8:
            mv = [5] + (rs.uniform(size=3 * len(r["seq"])) < 0.34).astype(int).tolist()
            mv[1] = 1
            tags["mv"] = mv
        r["tags"] = tags
    batch = bam.records_from_dicts(recs)
    start, end = 400, 2600
    want, wpos, wl, wr = read_matrix_oracle.read_alignment(recs, start, end, **kw)
    got, gpos, gl, gr = _device(batch, start, end, **kw)
    assert got.shape == want.shape
    assert np.array_equal(gpos, wpos)
    assert np.array_equal(got, want)
    assert gl == wl and gr == wr


@pytest.mark.gpu
def test_encoder_on_bam_file_with_sub_regions(tmp_path):
    """bam_to_sample through the native BAM reader with the region cut into sub-regions (row bookkeeping restarts in
    each, results joined on read identity) against the oracle driven the same way."""
    from medaka_b200 import common, features
    from tests import bamutil
    rs = np.random.RandomState(7)
    recs = synth.synth_reads(150, 2500, seed=11, mean_len=600)
    recs.sort(key=lambda r: r["pos"])
    for i, r in enumerate(recs):
        r["query_name"] = "q%d" % i
        r["ref"] = 0
        r["qual"] = rs.randint(1, 50, len(r["seq"])).tolist()
        r["tags"] = {}
    path = str(tmp_path / "reads.bam")
    bamutil.write_bam(path, [("ctg", 2500)], recs)
    region = common.Region("ctg", 100, 2400)
    enc = features.ReadAlignmentFeatureEncoder(include_dwells=False)
    samples = []
    for m, p in features.read_alignment_matrix(region, path, region_split=700, include_dwells=False):
        samples.append((m, p))
    want = []
    for sub in region.split(700, fixed_size=False):
        m, p, left, right = read_matrix_oracle.read_alignment(recs, sub.start, sub.end)
        want.append((m, p, (np.array([x.encode() for x in left], dtype="S"), np.array([x.encode() for x in right], dtype="S"))))
    want = features._join_read_matrix_chunks(want)
    assert len(samples) == len(want)
    for (m, p), (wm, wp) in zip(samples, want):
        assert np.array_equal(p, wp) and np.array_equal(m, wm)
    # the encoder's default sub-region size (100 kb) leaves this region in one piece: rows beyond the deepest column are
    # cut per sub-region (medaka/features.py:337-347), so the result legitimately depends on where the cuts fall
    out = enc.bam_to_sample(path, region)
    whole, wp, _, _ = read_matrix_oracle.read_alignment(recs, region.start, region.end)
    assert len(out) == 1 and out[0].features.dtype == np.int8
    assert np.array_equal(out[0].features, whole) and np.array_equal(out[0].positions, wp)
    assert np.array_equal(out[0].depth, np.count_nonzero(whole[..., 0], axis=-1))
