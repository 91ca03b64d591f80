"""The native BAM reader (csrc/bam_io.cu through medaka_b200.bam.BamFile) against files written from the SAM
specification by tests/bamutil.py: indexed and unindexed region fetches, records straddling BGZF members, CG-tag long
CIGARs, and the order of the read filters (src/medaka_bamiter.c:17-45).  Host code only: runs without a GPU."""
import os
import tempfile

import numpy as np
import pytest

from medaka_b200 import bam
from tests import bamutil


def _random_records(rs, n, ref_lens):
    recs = []
    for i in range(n):
        ref = int(rs.randint(0, len(ref_lens)))
        pos = int(rs.randint(0, ref_lens[ref] - 400))
        ops, qlen = [], 0
        if rs.uniform() < 0.3:
            s = int(rs.randint(1, 10)); ops.append("%dS" % s); qlen += s
        for _ in range(int(rs.randint(1, 6))):
            m = int(rs.randint(5, 60)); ops.append("%dM" % m); qlen += m
            kind = rs.uniform()
            if kind < 0.3:
                k = int(rs.randint(1, 4)); ops.append("%dI" % k); qlen += k
            elif kind < 0.6:
                ops.append("%dD" % int(rs.randint(1, 5)))
            elif kind < 0.65:
                ops.append("%dN" % int(rs.randint(5, 40)))
        m = int(rs.randint(5, 30)); ops.append("%dM" % m); qlen += m
        flag = int(rs.choice([0, 16, 0x100, 0x800, 0x4 | 0x10, 0x400], p=[0.4, 0.4, 0.05, 0.05, 0.05, 0.05]))
        recs.append(dict(ref=ref, pos=pos, cigar="".join(ops), seq="".join(rs.choice(list("ACGT"), qlen)), flag=flag,
                         mapq=int(rs.randint(0, 61)), query_name="read%04d" % i,
                         tags={"DT": "r10" if i % 3 else "r9"} if i % 7 else {}))
    recs.sort(key=lambda r: (r["ref"], r["pos"]))
    return recs


def _expect(recs, spans, ref, start, end, exclude, min_mapq):
    return [r["query_name"] for r, sp in zip(recs, spans)
            if r["ref"] == ref and r["pos"] < end and r["pos"] + sp > start and not (r["flag"] & exclude)
            and r["mapq"] >= min_mapq]


@pytest.mark.parametrize("indexed", [True, False])
def test_fetch_matches_brute_force(indexed):
    rs = np.random.RandomState(3)
    refs = [("ctgA", 40000), ("ctgB", 25000)]
    recs = _random_records(rs, 600, [l for _, l in refs])
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "t.bam")
        spans = bamutil.write_bam(path, refs, recs, member_size=777, with_index=indexed)
        bf = bam.BamFile(path, threads=3)
        assert bf.references == ["ctgA", "ctgB"] and bf.lengths == [40000, 25000] and bf.has_index == indexed
        for trial in range(40):
            ref = int(rs.randint(0, 2))
            a = int(rs.randint(0, refs[ref][1] - 10))
            b = int(min(refs[ref][1], a + rs.randint(1, 20000)))
            mq = int(rs.choice([0, 1, 30]))
            got = bf.fetch(refs[ref][0], a, b, with_names=True, min_mapq=mq)
            assert got.names == _expect(recs, spans, ref, a, b, bam.FILTER_FLAGS, mq), (ref, a, b, mq)
        # packed fields survive the trip
        got = bf.fetch("ctgA", 0, 40000, with_names=True, min_mapq=0, exclude_flags=0)
        by_name = {r["query_name"]: r for r in recs}
        assert len(got.names) == sum(1 for r in recs if r["ref"] == 0)
        ops = "MIDNSHP=X"
        nt = "=ACMGRSVTWYHKDBN"
        for i, nm in enumerate(got.names[:50]):
            r = by_name[nm]
            cig = "".join("%d%s" % (int(x) >> 4, ops[int(x) & 15]) for x in got.cigar[got.cigar_off[i]:got.cigar_off[i + 1]])
            assert cig == r["cigar"] and int(got.pos[i]) == r["pos"] and int(got.flag[i]) == r["flag"]
            sb = got.seq[got.seq_off[i]:got.seq_off[i + 1]]
            seq = "".join(nt[b >> 4] + nt[b & 15] for b in sb)[:int(got.l_seq[i])]
            assert seq == r["seq"]
        bf.close()


def test_long_cigar_is_resolved_from_the_cg_tag():
    """> 65535 operations: BAM stores <l_seq>S<ref_len>N and the real CIGAR in CG:B,I; htslib swaps it in when it
    reads the record, and so must we - otherwise the read widens the pileup but counts nothing."""
    recs = [dict(ref=0, pos=100, cigar="20M2I30M1D10M", seq="A" * 62, query_name="long", flag=0, mapq=60),
            dict(ref=0, pos=120, cigar="50M", seq="C" * 50, query_name="plain", flag=16, mapq=60)]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "t.bam")
        bamutil.write_bam(path, [("c", 1000)], recs, member_size=100, long_cigar_names=("long",))
        rb = bam.BamFile(path).fetch("c", 0, 1000, with_names=True)
        assert rb.names == ["long", "plain"]
        ops = "MIDNSHP=X"
        cig = "".join("%d%s" % (int(x) >> 4, ops[int(x) & 15]) for x in rb.cigar[rb.cigar_off[0]:rb.cigar_off[1]])
        assert cig == "20M2I30M1D10M"
        # the region test uses the real reference span (61), not the placeholder's
        assert bam.BamFile(path).fetch("c", 160, 162, with_names=True).names == ["long", "plain"]
        assert bam.BamFile(path).fetch("c", 161, 170, with_names=True).names == ["plain"]


def test_read_filters_run_in_the_reference_order():
    """medaka_bamiter.c:19-44: flags, then mapQ, then tag, then RG.  A secondary / low-mapQ read without a DT tag must
    be dropped silently, not abort the region with 'Datatype not found' (only reads that survive reach that test)."""
    mk = lambda name, flag=0, mapq=60, tags=None: dict(ref=0, pos=10, cigar="30M", seq="A" * 30, query_name=name,  # noqa
                                                       flag=flag, mapq=mapq, tags=tags or {})
    recs = [mk("ok9", tags={"DT": "r9", "RG": "g1", "XX": 5}), mk("ok10", 16, tags={"DT": "r10", "XX": 5}),
            mk("secondary_no_dt", 0x100), mk("lowq_no_dt", 0, 0), mk("wrongtag", tags={"DT": "r9", "XX": 4})]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "t.bam")
        bamutil.write_bam(path, [("c", 1000)], recs)
        bf = bam.BamFile(path)
        rb = bf.fetch("c", 0, 100, dtypes=("r9", "r10"), tag_name="XX", tag_value=5, with_names=True, min_mapq=1)
        assert rb.names == ["ok9", "ok10"] and rb.dtype.tolist() == [0, 1]
        with pytest.raises(ValueError):        # a PASSING read without the datatype tag still aborts, like the reference
            bf.fetch("c", 0, 100, dtypes=("r9", "r10"), with_names=True, min_mapq=0, exclude_flags=0)
        assert bf.fetch("c", 0, 100, read_group="g1", with_names=True).names == ["ok9"]
        assert bf.fetch("c", 0, 100, tag_name="XX", tag_value=5, keep_missing=True, with_names=True, min_mapq=1).names == \
            ["ok9", "ok10"]


def test_not_a_bam_is_an_error():
    from medaka_b200 import libmedaka
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "x.bam")
        with open(path, "wb") as fh:
            fh.write(b"hello world, definitely not BGZF")
        with pytest.raises(libmedaka.MedakaB200Error):
            bam.BamFile(path)
        with pytest.raises(libmedaka.MedakaB200Error):
            bam.BamFile(os.path.join(d, "missing.bam"))
