"""Pileup-counts featuriser: BAM reader + oracle on CPU fixtures; GPU kernel parity (bit-exact integers)."""
import os

import numpy as np
import pytest

from medaka_b200 import bam
from oracle import features_oracle, pileup_oracle, synth
from tests.test_oracle import EXPECTED_COUNTS, EXPECTED_POS, SIMPLE_CALLS


def _batch_from_fixture(g):
    return bam.RecordBatch(pos=g["pos"], flag=g["flag"], mapq=g["mapq"], dtype=g["dtype"], cigar=g["cigar"],
                           cigar_off=g["cigar_off"], seq=g["seq"], seq_off=g["seq_off"], l_seq=g["l_seq"],
                           names=None, tags=None)


def test_oracle_on_real_bam_slice(golden_dir):
    """The fixture was cut from the reference's test_reads.bam by tests/golden/make_pileup_golden.py, which also
    asserts the reference's regression numbers (86 294 columns, mean depth 18.696468) on the full region."""
    g = np.load(os.path.join(golden_dir, "pileup_real.npz"))
    c, p = pileup_oracle.pileup_counts_from_batch(_batch_from_fixture(g), int(g["start"]), int(g["end"]))
    assert np.array_equal(c, g["counts"]) and np.array_equal(p["major"], g["major"]) and np.array_equal(p["minor"], g["minor"])


def test_records_from_dicts_roundtrip():
    batch = bam.records_from_dicts(SIMPLE_CALLS)
    c, p = pileup_oracle.pileup_counts_from_batch(batch, 0, 8)
    assert np.array_equal(c, EXPECTED_COUNTS) and np.array_equal(p, EXPECTED_POS)


def test_bam_reader_roundtrip(tmp_path):
    """Write a tiny BAM by hand (BGZF blocks via zlib) and read it back with medaka_b200.bam."""
    import struct
    import zlib
    recs = SIMPLE_CALLS
    batch = bam.records_from_dicts(recs)
    body = b"BAM\x01" + struct.pack("<i", 0) + struct.pack("<i", 1) + struct.pack("<i", 4) + b"ref\x00" + struct.pack("<i", 8)
    for i, r in enumerate(recs):
        name = r["query_name"].encode() + b"\x00"
        cig = batch.cigar[batch.cigar_off[i]:batch.cigar_off[i + 1]].astype("<u4").tobytes()
        seq = batch.seq[batch.seq_off[i]:batch.seq_off[i + 1]].tobytes()
        qual = b"\xff" * len(r["seq"])
        tags = b"".join(k.encode() + (b"Z" + v.encode() + b"\x00" if isinstance(v, str) else b"C" + bytes([v]))
                        for k, v in r["tags"].items())
        core = struct.pack("<iiBBHHHiiii", 0, r["pos"], len(name), r["mapq"], 4680, len(cig) // 4, r["flag"],
                           len(r["seq"]), -1, -1, 0)
        rec = core + name + cig + seq + qual + tags
        body += struct.pack("<i", len(rec)) + rec

    def block(data):
        comp = zlib.compressobj(6, zlib.DEFLATED, -15)
        c = comp.compress(data) + comp.flush()
        bsize = len(c) + 25
        return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + c +
                struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))

    path = tmp_path / "simple.bam"
    path.write_bytes(block(body[:150]) + block(body[150:]) + block(b""))
    bf = bam.BamFile(str(path))
    assert bf.get_regions() == [("ref", 8)]
    rb = bf.fetch("ref", 0, 8, with_names=True)
    assert rb.names == [r["query_name"] for r in recs]
    c, p = pileup_oracle.pileup_counts_from_batch(rb, 0, 8)
    assert np.array_equal(c, EXPECTED_COUNTS) and np.array_equal(p, EXPECTED_POS)
    # tag filters done by the reader (medaka/test/test_counts.py:320-334)
    rb = bf.fetch("ref", 0, 8, tag_name="AA", tag_value=1)
    assert set(pileup_oracle.pileup_counts_from_batch(rb, 0, 8)[0].sum(axis=1)) == {2}
    rb = bf.fetch("ref", 0, 8, tag_name="AA", tag_value=1, keep_missing=True)
    assert set(pileup_oracle.pileup_counts_from_batch(rb, 0, 8)[0].sum(axis=1)) == {3}
    rb = bf.fetch("ref", 0, 8, dtypes=["r9", "r10"])
    assert rb.dtype.tolist() == [0, 0, 0, 1]


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_pileup_simple_golden():
    from medaka_b200 import features
    c, p = features.pileup_counts_from_batch(bam.records_from_dicts(SIMPLE_CALLS), 0, 8)
    assert np.array_equal(c, EXPECTED_COUNTS) and np.array_equal(p, EXPECTED_POS)      # test_counts.py:298-311
    c, p = features.pileup_counts_from_batch(bam.records_from_dicts(SIMPLE_CALLS, ["r9", "r10"]), 0, 8, num_dtypes=2)
    assert c.shape == (9, 20)                                                            # test_counts.py:348-353
    exp, _ = pileup_oracle.pileup_counts(SIMPLE_CALLS, 0, 8, dtypes=["r9", "r10"])
    assert np.array_equal(c, exp)
    # sub-regions, incl. one that starts inside the insertion column's neighbourhood
    for s, e in [(0, 3), (3, 8), (2, 5), (7, 8)]:
        c, p = features.pileup_counts_from_batch(bam.records_from_dicts(SIMPLE_CALLS), s, e)
        ec, ep = pileup_oracle.pileup_counts(SIMPLE_CALLS, s, e)
        assert np.array_equal(c, ec) and np.array_equal(p, ep), (s, e)


@pytest.mark.gpu
def test_gpu_pileup_real_bam_slice(golden_dir):
    from medaka_b200 import features
    g = np.load(os.path.join(golden_dir, "pileup_real.npz"))
    c, p = features.pileup_counts_from_batch(_batch_from_fixture(g), int(g["start"]), int(g["end"]))
    assert np.array_equal(c, g["counts"]) and np.array_equal(p["major"], g["major"]) and np.array_equal(p["minor"], g["minor"])


@pytest.mark.gpu
@pytest.mark.parametrize("seed,num_dtypes,p_skip,ins_after_skip",
                         [(0, 1, 0.0, False), (1, 1, 0.02, False), (2, 2, 0.0, False), (3, 1, 0.08, True), (4, 2, 0.05, True)])
def test_gpu_pileup_random_reads(seed, num_dtypes, p_skip, ins_after_skip):
    """(the last two cases put insertions right behind reference skips: they widen the column group of the skip's last
    position but are not counted, src/medaka_counts.c:259-263 vs :282)"""
    from medaka_b200 import features
    recs = synth.synth_reads(300, 4000, seed=seed, mean_len=800, p_skip=p_skip, num_dtypes=num_dtypes,
                             ins_after_skip=ins_after_skip)
    if ins_after_skip:
        assert any(__import__("re").search(r"\d+N\d+I", r["cigar"]) for r in recs)
    dtypes = ["dt%d" % k for k in range(num_dtypes)] if num_dtypes > 1 else None
    batch = bam.records_from_dicts(recs, dtypes)
    for s, e, mq in [(0, 4000, 1), (1000, 2500, 1), (500, 600, 20)]:
        c, p = features.pileup_counts_from_batch(batch, s, e, num_dtypes=num_dtypes, min_mapq=mq)
        ec, ep = pileup_oracle.pileup_counts(recs, s, e, dtypes=dtypes, min_mapq=mq)
        assert np.array_equal(p, ep), (s, e)
        assert np.array_equal(c, ec), (s, e)


@pytest.mark.gpu
def test_gpu_pileup_gaps_empty_and_overflow():
    from medaka_b200 import common, features
    # coverage gap -> two chunks (medaka/test/test_counts.py:229-243 behaviour); empty region -> no chunks
    reads = [dict(SIMPLE_CALLS[0], pos=0), dict(SIMPLE_CALLS[0], pos=20)]
    batch = bam.records_from_dicts(reads)
    c, p = features.pileup_counts_from_batch(batch, 0, 100)
    chunks = features._split_on_gaps(c, p)
    assert [len(x[1]) for x in chunks] == [8, 8]
    c, p = features.pileup_counts_from_batch(batch, 50, 60)
    assert len(c) == 0 and features._split_on_gaps(c, p) == []
    # more columns than 2*(end-start): the enlarge-and-retry path (medaka_counts.c:266-271)
    ins_read = dict(query_name="ins", pos=0, cigar="2M40I2M", seq="AC" + "G" * 40 + "TA", flag=0, mapq=60, tags={})
    c, p = features.pileup_counts_from_batch(bam.records_from_dicts([ins_read]), 0, 4)
    ec, ep = pileup_oracle.pileup_counts([ins_read], 0, 4)
    assert len(p) == 44 and np.array_equal(c, ec) and np.array_equal(p, ep)


@pytest.mark.gpu
def test_gpu_pileup_long_operations_and_many_reads():
    """Balance cases of the rewritten kernels: reads with > 32 CIGAR ops per warp step and multi-kilobase single
    operations (a 5 kb match, a 3 kb deletion, a 2 kb reference skip followed by an insertion), a region longer than
    one scan block, and depth in the hundreds."""
    from medaka_b200 import features
    rs = np.random.RandomState(5)
    seq = lambda n: "".join(rs.choice(list("ACGT"), n))  # noqa: E731
    recs = [dict(query_name="longM", pos=10, cigar="5000M", seq=seq(5000), flag=0, mapq=60, tags={}),
            dict(query_name="longD", pos=500, cigar="100M3000D100M", seq=seq(200), flag=16, mapq=60, tags={}),
            dict(query_name="skipI", pos=700, cigar="50M2000N4I60M", seq=seq(114), flag=0, mapq=60, tags={})]
    recs += synth.synth_reads(1500, 9000, seed=9, mean_len=2500, p_skip=0.01, ins_after_skip=True)
    recs.sort(key=lambda r: r["pos"])
    batch = bam.records_from_dicts(recs)
    for s, e in [(0, 9000), (2040, 4100), (4095, 4097)]:
        c, p = features.pileup_counts_from_batch(batch, s, e)
        ec, ep = pileup_oracle.pileup_counts(recs, s, e)
        assert np.array_equal(p, ep), (s, e)
        assert np.array_equal(c, ec), (s, e)
    assert int(c.sum()) > 0


@pytest.mark.gpu
def test_gpu_bam_to_sample_end_to_end(golden_dir):
    """BamFile-free path of CountsFeatureEncoder.bam_to_sample: GPU pileup -> gap split -> GPU normalise."""
    from medaka_b200 import common, features
    g = np.load(os.path.join(golden_dir, "pileup_real.npz"))
    batch = _batch_from_fixture(g)
    start, end = int(g["start"]), int(g["end"])
    enc = features.CountsFeatureEncoder(
        normalise="total",
        pileup_source=lambda region, b, e: features._split_on_gaps(
            *features.pileup_counts_from_batch(batch, region.start, region.end)))
    samples = enc.bam_to_sample(None, common.Region("utg000001l", start, end))
    assert len(samples) == 1
    pos = np.empty(len(g["major"]), dtype=[("major", "<i8"), ("minor", "<i8")])
    pos["major"], pos["minor"] = g["major"], g["minor"]
    ef, ed = features_oracle.post_process_pileup(g["counts"].copy(), pos, "total")
    assert np.array_equal(samples[0].features, ef) and np.array_equal(np.asarray(samples[0].depth), ed.astype(np.int64))


@pytest.mark.gpu
@pytest.mark.parametrize("normalise,sym,dtypes", [("total", False, None), ("fwd_rev", False, None), (None, True, None),
                                                  ("total", True, ("dt0", "dt1")), ("fwd_rev", False, ("dt0", "dt1"))])
def test_gpu_fused_pileup_features_match_two_step(normalise, sym, dtypes):
    """mdk_pileup_features (counts stay on the device) against oracle pileup + oracle post-processing, on reads with a
    coverage gap, consecutive insertions and two datatypes."""
    from medaka_b200 import bam, features
    recs = synth.synth_reads(140, 2400, seed=31, mean_len=300, num_dtypes=2 if dtypes else 1)
    recs = [r for r in recs if not (900 <= r["pos"] < 1100)]          # thin out a stretch: a gap in coverage
    for r in recs:
        if r["pos"] < 900:
            r["cigar"] = _clip_cigar(r["cigar"], 900 - r["pos"])
    batch = bam.records_from_dicts(recs, dtypes=dtypes)
    nd = len(dtypes) if dtypes else 1
    feats, depth, pos = features.pileup_features_from_batch(batch, 100, 2300, nd, 1, normalise, sym)
    ec, ep = pileup_oracle.pileup_counts(recs, 100, 2300, dtypes=dtypes)
    assert np.array_equal(pos, ep)
    assert (np.ediff1d(ep["major"]) > 1).any()
    # the reference normalises each gap-free piece on its own (features.py:125-134 then :871-935)
    cuts = np.where(np.ediff1d(ep["major"]) > 1)[0] + 1
    bounds = [0] + cuts.tolist() + [len(ep)]
    for a, b in zip(bounds[:-1], bounds[1:]):
        ef, ed = features_oracle.post_process_pileup(ec[a:b].copy(), ep[a:b], normalise, dtypes=dtypes or ("",), sym_indels=sym)
        assert np.array_equal(feats[a:b], ef)
        assert np.array_equal(depth[a:b], ed.astype(np.int64))


def _clip_cigar(cigar, max_ref):
    """Cut a CIGAR string so that it consumes at most max_ref reference bases (keeps it ending in a match)."""
    import re
    out, used = [], 0
    for n, op in re.findall(r"(\d+)([MIDNSHP=X])", cigar):
        n = int(n)
        if op in "MDN=X":
            if used + n > max_ref:
                n = max_ref - used
                if n > 0 and op in "M=X":
                    out.append("%d%s" % (n, op))
                break
            used += n
        out.append("%d%s" % (n, op))
    while out and out[-1][-1] not in "M=X":
        out.pop()
    return "".join(out) or "1M"


@pytest.mark.gpu
def test_gpu_encoder_fused_path_on_bam_file(tmp_path):
    """CountsFeatureEncoder.bam_to_sample on a BAM file (native reader -> fused device featuriser) equals the two-step
    path (pileup_counts -> _post_process_pileup) sample by sample."""
    from medaka_b200 import common, features
    from tests import bamutil
    recs = synth.synth_reads(120, 3000, seed=8, mean_len=400)
    recs = [r for r in recs if not (1400 <= r["pos"] < 1500)]
    recs.sort(key=lambda r: r["pos"])
    for r in recs:
        r["ref"] = 0
    path = str(tmp_path / "r.bam")
    bamutil.write_bam(path, [("ctg", 3000)], recs)
    region = common.Region("ctg", 0, 3000)
    for normalise in ("total", "fwd_rev", None):
        enc = features.CountsFeatureEncoder(normalise=normalise, sym_indels=normalise == "total")
        fused = enc.bam_to_sample(path, region)
        two = [enc._post_process_pileup(c, p, region) for c, p in enc._pileup_function(region, path)]
        assert len(fused) == len(two) >= 1
        for a, b in zip(fused, two):
            assert np.array_equal(a.positions, b.positions) and np.array_equal(a.features, b.features)
            assert np.array_equal(np.asarray(a.depth), np.asarray(b.depth))
