"""Test helper: write small BAM files (BGZF members of a chosen size) and their .bai index from dict records, straight
from the SAM specification (sections 4.1, 4.2, 5.2) - independent of the reader under test (csrc/bam_io.cu)."""
import re
import struct
import zlib

import numpy as np

_OPS = {c: i for i, c in enumerate("MIDNSHP=X")}
_NT = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
_CONSUMES_REF = set("MDN=X")


def _aux(tags):
    out = b""
    for k, v in (tags or {}).items():
        if isinstance(v, int):
            out += k.encode() + b"i" + struct.pack("<i", v)
        elif isinstance(v, str):
            out += k.encode() + b"Z" + v.encode() + b"\x00"
        elif isinstance(v, (list, tuple, np.ndarray)):     # B,I array
            out += k.encode() + b"BI" + struct.pack("<I", len(v)) + b"".join(struct.pack("<I", int(x)) for x in v)
        else:
            raise TypeError(v)
    return out


def reg2bin(beg, end):
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def encode_record(rec, tid, long_cigar=False):
    ops = [(int(n) << 4) | _OPS[o] for n, o in re.findall(r"(\d+)([MIDNSHP=X])", rec["cigar"])]
    ref_len = sum(int(n) for n, o in re.findall(r"(\d+)([MIDNSHP=X])", rec["cigar"]) if o in _CONSUMES_REF)
    seq = rec["seq"].upper()
    tags = dict(rec.get("tags") or {})
    if long_cigar:        # SAM spec 4.2.2: placeholder <l_seq>S<ref_len>N + the real CIGAR in CG:B,I
        tags["CG"] = list(ops)
        ops = [(len(seq) << 4) | _OPS["S"], (ref_len << 4) | _OPS["N"]]
    name = (rec.get("query_name") or "r").encode() + b"\x00"
    nib = [_NT.get(c, 15) for c in seq] + ([0] if len(seq) % 2 else [])
    packed = bytes((nib[i] << 4) | nib[i + 1] for i in range(0, len(nib), 2))
    body = struct.pack("<iiBBHHHiiii", tid, rec["pos"], len(name), rec.get("mapq", 60),
                       reg2bin(rec["pos"], rec["pos"] + max(ref_len, 1)), len(ops), rec.get("flag", 0), len(seq), -1, -1, 0)
    qual = bytes(rec["qual"]) if rec.get("qual") is not None else b"\xff" * len(seq)
    body += name + b"".join(struct.pack("<I", o) for o in ops) + packed + qual + _aux(tags)
    return struct.pack("<i", len(body)) + body, ref_len


def _member(data):
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    bsize = len(comp) + 25
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + comp +
            struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))


def write_bam(path, refs, records, member_size=600, long_cigar_names=(), with_index=True):
    """refs: [(name, length)]; records: dicts with 'ref' (index), 'pos', 'cigar', 'seq', optional flag/mapq/tags/
    query_name; must be sorted by (ref, pos).  Returns the list of (ref_len) per record."""
    header = b"BAM\x01" + struct.pack("<i", 0) + struct.pack("<i", len(refs))
    for name, length in refs:
        header += struct.pack("<i", len(name) + 1) + name.encode() + b"\x00" + struct.pack("<i", length)
    stream = header
    rec_u = []          # (uncompressed start, end) of each record
    spans = []
    for rec in records:
        enc, ref_len = encode_record(rec, rec["ref"], rec.get("query_name") in long_cigar_names)
        rec_u.append((len(stream), len(stream) + len(enc)))
        spans.append(ref_len)
        stream += enc
    # cut into members; remember (file offset, uncompressed offset) of each member start
    out, starts, u = b"", [], 0
    while u < len(stream):
        starts.append((len(out), u))
        out += _member(stream[u:u + member_size])
        u += member_size
    eof_at = len(out)
    out += _member(b"")
    with open(path, "wb") as fh:
        fh.write(out)

    def voff(upos):
        k = max(i for i, (_, us) in enumerate(starts) if us <= upos)
        if upos == len(stream):
            return eof_at << 16
        return (starts[k][0] << 16) | (upos - starts[k][1])

    if with_index:
        bai = b"BAI\x01" + struct.pack("<i", len(refs))
        for tid in range(len(refs)):
            bins, linear = {}, {}
            for rec, (u0, u1), ref_len in zip(records, rec_u, spans):
                if rec["ref"] != tid:
                    continue
                beg, end = rec["pos"], rec["pos"] + max(ref_len, 1)
                bins.setdefault(reg2bin(beg, end), []).append((voff(u0), voff(u1)))
                for w in range(beg >> 14, ((end - 1) >> 14) + 1):
                    linear[w] = min(linear.get(w, 1 << 62), voff(u0))
            bai += struct.pack("<i", len(bins))
            for b, chunks in sorted(bins.items()):
                bai += struct.pack("<Ii", b, len(chunks)) + b"".join(struct.pack("<QQ", a, e) for a, e in chunks)
            n_intv = (max(linear) + 1) if linear else 0
            bai += struct.pack("<i", n_intv)
            last = 0
            for w in range(n_intv):
                last = linear.get(w, last)
                bai += struct.pack("<Q", last)
        with open(path + ".bai", "wb") as fh:
            fh.write(bai)
    return spans
