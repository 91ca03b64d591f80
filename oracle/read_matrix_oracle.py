"""Oracle for the read-level featuriser: calculate_read_alignment restated over plain records.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pure-Python loops: small cases only.

Follows src/medaka_read_matrix.c:277-615 statement by statement (row bookkeeping included: the read array, the
qname -> row hash that is never pruned, the buffer growth rule :359-371, the `read_i >= buffer_reads` skip :462) with
the pileup fields restated as in oracle/pileup_oracle.py (htslib 1.14 is not in the tree).  Dwell times follow
calculate_dwells (:154-213).  The Python wrapper's clip (medaka/features.py:309-310, `np.maximum(counts, 0)`) is
`clip_to_zero`.

A record is a dict: {'query_name', 'pos', 'cigar', 'seq', 'qual' (list of ints, or None = missing -> 0xff per base),
'flag', 'mapq', 'tags': {'HP': int, 'DT': str, 'mv': [stride, m0, m1, ...]}}.
"""
import numpy as np

from oracle import pileup_oracle

BASE_FEATLEN = 4      # src/medaka_read_matrix.h:37
DEL_VAL = 5           # :38
MIN_GAP = 5           # medaka_read_matrix.c:329
# src/medaka_read_matrix.h:41-46 : 4-bit IUPAC code -> 1..4 (ACGT), -1 otherwise
NUM2COUNTBASE_SYMM = [-1, 1, 2, -1, 3, -1, -1, -1, 4, -1, -1, -1, -1, -1, -1, -1]


def _i8(v):
    v = int(v) & 0xFF
    return v - 256 if v > 127 else v


def aligned_ref_len(rec):
    """aligned_ref_pos_from_cigar (:258-273): M, D, =, X only - reference skips are NOT counted."""
    return sum(int(n) for n, op in pileup_oracle._CIGAR_RE.findall(rec["cigar"]) if op in "MD=X")


def calculate_dwells(rec):
    """calculate_dwells (:154-213); None when there is no move table or it does not fit the sequence."""
    mv = rec.get("tags", {}).get("mv")
    if mv is None:
        return None
    length = len(rec["seq"])
    mv_len = len(mv)
    out = [0] * length
    qpos = 0
    if rec.get("flag", 0) & 0x10:
        dwell = 0
        for i in range(mv_len - 1, 0, -1):
            dwell += 1
            if mv[i] == 1:
                if qpos >= length:
                    return None
                out[qpos] = _i8(min(dwell, 127))
                qpos += 1
                dwell = 0
    else:
        dwell = 1
        for i in range(2, mv_len):
            if mv[i] == 1:
                if qpos >= length:
                    return None
                out[qpos] = _i8(min(dwell, 127))
                qpos += 1
                dwell = 0
            dwell += 1
        if qpos >= length:
            return None          # (the C writes one past the array here; a malformed table, as in the reference's test)
        out[qpos] = _i8(min(dwell, 127))
    return out


def read_alignment(records, start, end, dtypes=None, min_mapq=1, tag_name=None, tag_value=None, keep_missing=False,
                   read_group=None, row_per_read=False, include_dwells=False, include_haplotype=False, max_reads=100,
                   clip_to_zero=True):
    """calculate_read_alignment for one contig, region [start, end).

    Returns (matrix int8 [n_pos, n_reads, featlen], positions [('major','minor')], read_ids_left, read_ids_right)."""
    num_dtypes = 1 if not dtypes or len(dtypes) == 1 else len(dtypes)
    featlen = BASE_FEATLEN + (1 if include_dwells else 0) + (1 if include_haplotype else 0) + (1 if num_dtypes > 1 else 0)
    cols = {}
    covered = set()
    for rec in records:
        if not pileup_oracle.read_passes(rec, min_mapq, tag_name, tag_value, keep_missing, read_group):
            continue
        for pos, is_del, is_refskip, indel, qpos in pileup_oracle._resolve_cigar(rec):
            covered.add(pos)
            if start <= pos < end:
                cols.setdefault(pos, []).append((rec, is_del, is_refskip, indel, qpos))
    # `pos` when the column loop ends (:337-341): the first pileup position at or behind `end` (the `break`), or the last
    # position the iterator returned
    beyond = [p for p in covered if p >= end]
    final_pos = min(beyond) if beyond else (max(covered) if covered else 0)
    buffer_reads = min(max_reads, 100)                                   # :326
    columns = []            # per emitted column: dict row -> feature vector
    major, minor = [], []
    read_array = []         # the kvec of Read structs
    read_map = {}           # qname -> row (never pruned)
    left = {}
    max_n_reads = 0
    n_pos = 0
    for pos in sorted(cols):
        plp = cols[pos]
        n_pos += 1
        n_plp = len(plp)
        max_ins = max([p[3] for p in plp if p[3] > 0], default=0)        # :346-351
        if n_plp > max_n_reads:
            max_n_reads = n_plp
        if buffer_reads < max_reads and max_n_reads + (n_plp if row_per_read else 0) > buffer_reads:   # :359-371
            buffer_reads = min(max_reads, max(max_n_reads + (n_plp if row_per_read else 0), 2 * buffer_reads))
        block = [dict() for _ in range(max_ins + 1)]
        for rec, is_del, is_refskip, indel, qpos in plp:
            if is_refskip:                                               # :381
                continue
            qname = rec["query_name"]
            read_i = read_map.get(qname, -1)
            if read_i == -1:                                             # a new read (:389-459)
                dtype = 0
                if num_dtypes > 1:
                    dtype = list(dtypes).index(rec["tags"]["DT"])
                tags = rec.get("tags", {})
                read = dict(rec=rec, strand=-1 if rec.get("flag", 0) & 0x10 else 1, mq=_i8(rec.get("mapq", 60)),
                            haplotype=int(tags.get("HP", 0)) & 0xFF, dtype=dtype,
                            ref_end=rec["pos"] + aligned_ref_len(rec),
                            dwells=calculate_dwells(rec) if include_dwells else None)
                array_size = len(read_array)
                placed = False
                if not row_per_read:
                    for read_i in range(array_size):
                        if pos >= read_array[read_i]["ref_end"] + MIN_GAP:
                            read_array[read_i] = read
                            read_map[qname] = read_i
                            placed = True
                            break
                    else:
                        read_i = array_size
                else:
                    read_i = array_size
                    if array_size > max_n_reads:
                        max_n_reads = array_size
                if not placed and read_i == array_size:
                    if read_i < buffer_reads:
                        read_array.append(read)
                    read_map[qname] = read_i
            if read_i >= buffer_reads:                                   # :462
                continue
            if read_i >= len(read_array):
                # the reference would read a struct that was never pushed (undefined behaviour); cannot be restated
                raise RuntimeError("reference behaviour undefined: row %d beyond the read array" % read_i)
            read = read_array[read_i]
            if n_pos == 1:
                left[read_i] = read["rec"]["query_name"]
            src = read["rec"]
            quals = src.get("qual")

            def cell(base, qual, dwell):
                v = [_i8(base), _i8(qual), read["strand"], read["mq"]]
                if include_dwells:
                    v.append(_i8(dwell))
                if include_haplotype:
                    v.append(_i8(read["haplotype"]))
                if num_dtypes > 1:
                    v.append(_i8(read["dtype"]))
                return v
            min_minor = 0
            max_minor = indel if indel > 0 else 0
            if is_del:                                                   # :473-494
                block[0][read_i] = cell(DEL_VAL, -1, -1)
                min_minor = 1
            off = 0
            mn = min_minor
            while mn <= max_minor:                                       # :498-526
                q = qpos + off
                code = pileup_oracle.SEQ_NT16.get(src["seq"][q].upper(), 15)
                base_i = NUM2COUNTBASE_SYMM[code]
                qv = 0xFF if quals is None else quals[q]
                dw = read["dwells"][q] if (include_dwells and read["dwells"] is not None) else 0   # untouched cell
                block[mn][read_i] = cell(base_i, qv, dw)
                mn += 1
                off += 1
            while mn <= max_ins:                                         # :527-553
                block[mn][read_i] = cell(DEL_VAL, -1, -1)
                mn += 1
        for i in range(max_ins + 1):
            columns.append(block[i])
            major.append(pos)
            minor.append(i)
        n_pos += max_ins
    ids_left, ids_right = [], []
    nleft = nright = 0
    for r, read in enumerate(read_array):                                # :559-575
        if read["ref_end"] >= final_pos:
            ids_right.append(read["rec"]["query_name"])
        else:
            nright += 1
            ids_right.append("__blank_%d" % nright)
        if r in left:
            ids_left.append(left[r])
        else:
            nleft += 1
            ids_left.append("__blank_%d" % nleft)
    n_reads = min(max_reads, len(read_array) if row_per_read else max_n_reads)      # :577-583
    mat = np.zeros((len(columns), max(n_reads, 0), featlen), dtype=np.int8)
    for c, col in enumerate(columns):
        for r, v in col.items():
            if r < n_reads:
                mat[c, r] = v
    if clip_to_zero:
        mat = np.maximum(mat, 0)
    positions = np.empty(len(columns), dtype=[("major", "<i8"), ("minor", "<i8")])
    positions["major"] = major
    positions["minor"] = minor
    pad = lambda ids: (ids + [""] * n_reads)[:n_reads]        # rows beyond the read array carry NULL ids (b"" in Python)
    return mat, positions, pad(ids_left), pad(ids_right)
