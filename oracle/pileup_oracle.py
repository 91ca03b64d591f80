"""Oracle for the pileup-counts featuriser: calculate_pileup restated over plain records.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pure-Python loops: small cases only.

Follows src/medaka_counts.c:199-372 (column loop :251-361, table
src/medaka_counts.h:25-30) and the read filter src/medaka_bamiter.c:17-45.
The per-column ``bam_pileup1_t`` fields (is_del, is_refskip, indel, qpos) come from
htslib 1.14 (build.py:10, NOT in the tree); ``_resolve_cigar`` restates its published
``resolve_cigar2`` behaviour: the "peek the next operation" rule sets indel at the
last reference base of an op (M or D alike); for a deletion qpos is the index of the
next query base; leading insertions/soft clips attach to no column.
num_homop == 1 only (q-score stratification / Weibull summation belong to the legacy
RLE models and are out of scope, SURVEY.md 2.1 ``rle.py``).

A record is a dict: {'query_name', 'pos' (0-based reference_start), 'cigar' (string),
'seq', 'flag', 'mapq', 'tags': {...}}.
"""
import re

import numpy as np

# src/medaka_counts.h:25-30 : 4-bit IUPAC code (+16 if reverse) -> index in 'acgtACGTdD'
NUM2COUNTBASE = [
    -1, 4, 5, -1, 6, -1, -1, -1,
    7, -1, -1, -1, -1, -1, -1, -1,
    -1, 0, 1, -1, 2, -1, -1, -1,
    3, -1, -1, -1, -1, -1, -1, -1,
]
FEATLEN, FWD_DEL, REV_DEL = 10, 9, 8
SEQ_NT16 = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
_FILTER_FLAGS = 0x4 | 0x100 | 0x800 | 0x200 | 0x400  # UNMAP|SECONDARY|SUPPLEMENTARY|QCFAIL|DUP
_CIGAR_RE = re.compile(r"(\d+)([MIDNSHP=X])")


def read_passes(rec, min_mapq=1, tag_name=None, tag_value=None, keep_missing=False,
                read_group=None):
    """src/medaka_bamiter.c:17-45, including the early accept on keep_missing."""
    if rec.get("flag", 0) & _FILTER_FLAGS:
        return False
    if int(rec.get("mapq", 60)) < min_mapq:
        return False
    tags = rec.get("tags", {})
    if tag_name:
        if tag_name not in tags:
            return bool(keep_missing)     # 'break' at bamiter.c:30 skips the RG filter too
        if not isinstance(tags[tag_name], int) or tags[tag_name] != tag_value:
            return False
    if read_group is not None:
        if tags.get("RG") != read_group:
            return False
    return True


def _resolve_cigar(rec):
    """Yield (ref_pos, is_del, is_refskip, indel, qpos) for every covered reference position."""
    ops = [(op, int(n)) for n, op in _CIGAR_RE.findall(rec["cigar"])]
    x, y = rec["pos"], 0
    for k, (op, l) in enumerate(ops):
        if op in "M=X" or op in "DN":
            for off in range(l):
                pos = x + off
                indel = 0
                if off == l - 1 and k + 1 < len(ops):
                    op2, l2 = ops[k + 1]
                    if op2 == "D" and op != "D":
                        indel = -l2
                        for op3, l3 in ops[k + 2:]:
                            if op3 == "D":
                                indel -= l3
                            else:
                                break
                    elif op2 == "I":
                        indel = l2
                        for op3, l3 in ops[k + 2:]:
                            if op3 == "I":
                                indel += l3
                            elif op3 != "P":
                                break
                if op in "M=X":
                    yield pos, False, False, indel, y + off
                else:
                    yield pos, True, op == "N", indel, y
            x += l
            if op in "M=X":
                y += l
        elif op in "IS":
            y += l
        # H, P consume nothing


def pileup_counts(records, start, end, dtypes=None, min_mapq=1, tag_name=None,
                  tag_value=None, keep_missing=False, read_group=None):
    """calculate_pileup (src/medaka_counts.c:199-372) for one contig, region [start, end).

    Returns (counts uint64 [n_cols, 10*num_dtypes], positions [('major','minor')]).
    Only reference positions covered by >= 1 passing read produce columns.
    """
    num_dtypes = 1 if not dtypes or len(dtypes) == 1 else len(dtypes)
    F = FEATLEN * num_dtypes
    cols = {}   # ref_pos -> list of (rec, is_del, is_refskip, indel, qpos)
    for rec in records:
        if not read_passes(rec, min_mapq, tag_name, tag_value, keep_missing, read_group):
            continue
        for pos, is_del, is_refskip, indel, qpos in _resolve_cigar(rec):
            if start <= pos < end:
                cols.setdefault(pos, []).append((rec, is_del, is_refskip, indel, qpos))
    rows, major, minor = [], [], []
    for pos in sorted(cols):
        plp = cols[pos]
        max_ins = max([p[3] for p in plp if p[3] > 0], default=0)       # .c:259-263
        block = np.zeros((max_ins + 1, F), dtype=np.uint64)
        for rec, is_del, is_refskip, indel, qpos in plp:
            if is_refskip:                                              # .c:282
                continue
            dtype = 0
            if num_dtypes > 1:                                          # .c:285-311
                dtype = list(dtypes).index(rec["tags"]["DT"])
            rev = bool(rec.get("flag", 0) & 0x10)
            min_minor = 0
            max_minor = indel if indel > 0 else 0
            if is_del:                                                  # .c:315-322
                block[0, FEATLEN * dtype + (REV_DEL if rev else FWD_DEL)] += 1
                min_minor = 1
            off = 0
            for mn in range(min_minor, max_minor + 1):                  # .c:324-357
                base_j = SEQ_NT16.get(rec["seq"][qpos + off].upper(), 15)
                if rev:
                    base_j += 16
                base_i = NUM2COUNTBASE[base_j]
                if base_i != -1:
                    block[mn, FEATLEN * dtype + base_i] += 1
                off += 1
        for i in range(max_ins + 1):
            rows.append(block[i])
            major.append(pos)
            minor.append(i)
    positions = np.empty(len(rows), dtype=[("major", "<i8"), ("minor", "<i8")])
    positions["major"] = major
    positions["minor"] = minor
    counts = np.stack(rows) if rows else np.zeros((0, F), dtype=np.uint64)
    return counts, positions


def records_from_batch(batch):
    """RecordBatch (medaka_b200.bam: BAM-packed CIGAR ops and 4-bit sequence) -> the dict records used above."""
    ops = "MIDNSHP=X"
    nt = "=ACMGRSVTWYHKDBN"
    recs = []
    for i in range(len(batch.pos)):
        c = batch.cigar[batch.cigar_off[i]:batch.cigar_off[i + 1]]
        cig = "".join("%d%s" % (int(x) >> 4, ops[int(x) & 15]) for x in c)
        sb = np.asarray(batch.seq[batch.seq_off[i]:batch.seq_off[i + 1]])
        nib = np.empty(2 * len(sb), dtype=np.uint8)
        nib[0::2] = sb >> 4
        nib[1::2] = sb & 15
        seq = "".join(nt[x] for x in nib[:int(batch.l_seq[i])])
        tags = {"DT": None}
        recs.append(dict(pos=int(batch.pos[i]), cigar=cig, seq=seq, flag=int(batch.flag[i]),
                         mapq=int(batch.mapq[i]), tags=tags, _dtype=int(batch.dtype[i])))
    return recs


def pileup_counts_from_batch(batch, start, end, num_dtypes=1, min_mapq=1):
    """calculate_pileup over a RecordBatch; dtype indices come pre-resolved in ``batch.dtype``."""
    recs = records_from_batch(batch)
    dtypes = None
    if num_dtypes > 1:
        dtypes = ["dt%d" % k for k in range(num_dtypes)]
        for r in recs:
            r["tags"]["DT"] = dtypes[r["_dtype"]]
    return pileup_counts(recs, start, end, dtypes=dtypes, min_mapq=min_mapq)
