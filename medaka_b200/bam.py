"""A small BAM reader (BGZF inflate + record parsing) feeding the GPU pileup-counts featuriser.

The reference reads alignments through htslib (``bam_itr_querys`` / ``bam_mplp_auto``,
src/medaka_counts.c:233-251), which is not part of its tree and not available here; the
engine does the per-base work on the GPU instead, so the host only has to (1) inflate the
BGZF blocks and (2) slice out, per alignment record, the fields the featuriser consumes in
BAM's own packed encodings: 32-bit CIGAR ops (``len << 4 | op``) and 4-bit sequence codes.
File format per the SAM/BAM specification (sections 4.1 BGZF, 4.2 BAM).

``BamFile`` inflates the whole file once (zlib releases the GIL, so blocks inflate in a
thread pool) and keeps numpy arrays of the fixed-width fields; ``fetch(ref_name, start,
end)`` then returns the records overlapping a region (what ``bam_itr_querys`` yields) as a
``RecordBatch`` of flat arrays ready for ``mdk_pileup_counts``.
"""
import collections
import concurrent.futures
import struct
import zlib

import numpy as np

# flags the pileup never sees (src/medaka_bamiter.c:19): UNMAP | SECONDARY | QCFAIL | DUP | SUPPLEMENTARY
FILTER_FLAGS = 0x4 | 0x100 | 0x200 | 0x400 | 0x800
_CONSUMES_REF = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0], dtype=np.int64)    # M I D N S H P = X
_CONSUMES_QRY = np.array([1, 1, 0, 0, 1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0], dtype=np.int64)

RecordBatch = collections.namedtuple(
    "RecordBatch",
    ["pos", "flag", "mapq", "dtype", "cigar", "cigar_off", "seq", "seq_off", "l_seq", "names", "tags"])


def _bgzf_blocks(buf):
    """Yield (compressed payload start, payload length, uncompressed size) for every BGZF block."""
    off, n = 0, len(buf)
    while off < n:
        if buf[off:off + 4] != b"\x1f\x8b\x08\x04":
            raise ValueError("not a BGZF block at offset {}".format(off))
        xlen = struct.unpack_from("<H", buf, off + 10)[0]
        bsize = None
        x = off + 12
        while x < off + 12 + xlen:
            si1, si2, slen = buf[x], buf[x + 1], struct.unpack_from("<H", buf, x + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", buf, x + 4)[0]
            x += 4 + slen
        if bsize is None:
            raise ValueError("BGZF block without BC subfield")
        cstart = off + 12 + xlen
        clen = bsize - xlen - 19
        isize = struct.unpack_from("<I", buf, off + bsize - 3)[0]
        yield cstart, clen, isize
        off += bsize + 1


def _inflate(args):
    buf, cstart, clen = args
    return zlib.decompress(buf[cstart:cstart + clen], -15)


def bgzf_decompress(buf, threads=4):
    blocks = [(buf, c, l) for c, l, isize in _bgzf_blocks(buf) if isize > 0]
    if threads > 1 and len(blocks) > 8:
        with concurrent.futures.ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(_inflate, blocks, chunksize=16))
    else:
        parts = [_inflate(b) for b in blocks]
    return b"".join(parts)


def _parse_tags(raw):
    """Aux fields -> dict (only the scalar / string types the read filters look at; arrays are skipped)."""
    tags, i, n = {}, 0, len(raw)
    sizes = {"c": ("<b", 1), "C": ("<B", 1), "s": ("<h", 2), "S": ("<H", 2), "i": ("<i", 4), "I": ("<I", 4),
             "f": ("<f", 4)}
    while i + 3 <= n:
        tag = raw[i:i + 2].decode("latin1")
        typ = chr(raw[i + 2])
        i += 3
        if typ in sizes:
            fmt, sz = sizes[typ]
            tags[tag] = struct.unpack_from(fmt, raw, i)[0]
            i += sz
        elif typ == "A":
            tags[tag] = chr(raw[i])
            i += 1
        elif typ in "ZH":
            j = raw.index(b"\x00", i)
            tags[tag] = raw[i:j].decode("latin1")
            i = j + 1
        elif typ == "B":
            sub = chr(raw[i])
            cnt = struct.unpack_from("<I", raw, i + 1)[0]
            i += 5 + cnt * {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}[sub]
        else:
            break
    return tags


def _passes_tag_filters(tags, tag_name, tag_value, keep_missing, read_group):
    """Tag part of the read filter (src/medaka_bamiter.c:24-44): a read lacking ``tag_name`` is accepted or
    rejected on the spot according to ``keep_missing`` (the accept path skips the RG test, :28-33)."""
    if tag_name:
        if tag_name not in tags:
            return bool(keep_missing)
        if not isinstance(tags[tag_name], int) or tags[tag_name] != tag_value:
            return False
    if read_group is not None and tags.get("RG") != read_group:
        return False
    return True


class BamFile(object):
    """All alignment records of a BAM file, inflated and indexed in memory."""

    def __init__(self, path, threads=4):
        with open(path, "rb") as fh:
            data = bgzf_decompress(fh.read(), threads)
        self.data = data
        if data[:4] != b"BAM\x01":
            raise ValueError("{} is not a BAM file".format(path))
        l_text = struct.unpack_from("<i", data, 4)[0]
        off = 8 + l_text
        n_ref = struct.unpack_from("<i", data, off)[0]
        off += 4
        self.references, self.lengths = [], []
        for _ in range(n_ref):
            l_name = struct.unpack_from("<i", data, off)[0]
            self.references.append(data[off + 4:off + 4 + l_name - 1].decode())
            self.lengths.append(struct.unpack_from("<i", data, off + 4 + l_name)[0])
            off += 8 + l_name
        # record offsets (sequential walk over block_size fields)
        offs = []
        n = len(data)
        while off + 4 <= n:
            bs = struct.unpack_from("<i", data, off)[0]
            offs.append(off)
            off += 4 + bs
        self.rec_off = np.array(offs, dtype=np.int64)
        raw = np.frombuffer(data, dtype=np.uint8)
        self._raw = raw

        def field(delta, dtype):
            width = np.dtype(dtype).itemsize
            idx = self.rec_off[:, None] + delta + np.arange(width)[None, :]
            return raw[idx].copy().view(dtype).reshape(-1)

        self.block_size = field(0, "<i4")
        self.ref_id = field(4, "<i4")
        self.pos = field(8, "<i4")
        self.l_read_name = field(12, "u1").astype(np.int64)
        self.mapq = field(13, "u1")
        self.n_cigar = field(16, "<u2").astype(np.int64)
        self.flag = field(18, "<u2")
        self.l_seq = field(20, "<i4").astype(np.int64)
        self.cigar_start = self.rec_off + 36 + self.l_read_name
        self.seq_start = self.cigar_start + 4 * self.n_cigar
        self.qual_start = self.seq_start + (self.l_seq + 1) // 2
        self.tag_start = self.qual_start + self.l_seq
        self.rec_end = self.rec_off + 4 + self.block_size
        # reference end of every record from its CIGAR (vectorised)
        total_ops = int(self.n_cigar.sum())
        op_rec = np.repeat(np.arange(len(offs)), self.n_cigar)
        first = np.cumsum(self.n_cigar) - self.n_cigar
        op_idx = np.arange(total_ops) - np.repeat(first, self.n_cigar)
        op_addr = self.cigar_start[op_rec] + 4 * op_idx
        ops = raw[op_addr[:, None] + np.arange(4)[None, :]].copy().view("<u4").reshape(-1)
        ref_len = np.zeros(len(offs), dtype=np.int64)
        np.add.at(ref_len, op_rec, (ops >> 4).astype(np.int64) * _CONSUMES_REF[ops & 0xF])
        self.end = self.pos.astype(np.int64) + ref_len
        self._ops = ops
        self._op_first = first

    def get_regions(self):
        """(name, length) of every reference sequence - what get_bam_regions needs (medaka/common.py:762-790)."""
        return list(zip(self.references, self.lengths))

    def name(self, i):
        s = int(self.rec_off[i]) + 36
        return self.data[s:s + int(self.l_read_name[i]) - 1].decode()

    def tags(self, i):
        return _parse_tags(self.data[int(self.tag_start[i]):int(self.rec_end[i])])

    def fetch(self, ref_name, start, end, dtypes=None, tag_name=None, tag_value=None, keep_missing=False,
              read_group=None, with_names=False):
        """Records overlapping [start, end) on ref_name, with the tag-based read filters of
        src/medaka_bamiter.c:24-44 applied on the host (flag / mapq filters run on the device)."""
        tid = self.references.index(ref_name)
        sel = np.flatnonzero((self.ref_id == tid) & (self.pos < end) & (self.end > start) &
                             ((self.flag & 0x4) == 0))
        need_tags = bool(tag_name) or read_group is not None or (dtypes is not None and len(dtypes) > 1)
        dtype = np.zeros(len(sel), dtype=np.uint8)
        tags_out = None
        if need_tags:
            keep = np.ones(len(sel), dtype=bool)
            tags_out = []
            for k, i in enumerate(sel):
                tg = self.tags(i)
                tags_out.append(tg)
                keep[k] = _passes_tag_filters(tg, tag_name, tag_value, keep_missing, read_group)
                if keep[k] and dtypes is not None and len(dtypes) > 1:
                    if tg.get("DT") not in dtypes:
                        raise ValueError("Datatype not found for {}.".format(self.name(i)))
                    dtype[k] = list(dtypes).index(tg["DT"])
            sel, dtype = sel[keep], dtype[keep]
            tags_out = [t for t, k in zip(tags_out, keep) if k]
        n_cig = self.n_cigar[sel]
        cigar_off = np.zeros(len(sel) + 1, dtype=np.int64)
        np.cumsum(n_cig, out=cigar_off[1:])
        op_rec = np.repeat(np.arange(len(sel)), n_cig)
        op_idx = np.arange(int(cigar_off[-1])) - np.repeat(cigar_off[:-1], n_cig)
        cigar = self._ops[self._op_first[sel][op_rec] + op_idx] if len(sel) else np.zeros(0, dtype="<u4")
        seq_bytes = (self.l_seq[sel] + 1) // 2
        seq_off = np.zeros(len(sel) + 1, dtype=np.int64)
        np.cumsum(seq_bytes, out=seq_off[1:])
        b_rec = np.repeat(np.arange(len(sel)), seq_bytes)
        b_idx = np.arange(int(seq_off[-1])) - np.repeat(seq_off[:-1], seq_bytes)
        seq = self._raw[self.seq_start[sel][b_rec] + b_idx] if len(sel) else np.zeros(0, dtype=np.uint8)
        names = [self.name(i) for i in sel] if with_names else None
        return RecordBatch(pos=self.pos[sel].astype(np.int32), flag=self.flag[sel].astype(np.uint16),
                           mapq=self.mapq[sel].astype(np.uint8), dtype=dtype,
                           cigar=np.ascontiguousarray(cigar, dtype=np.uint32), cigar_off=cigar_off,
                           seq=np.ascontiguousarray(seq, dtype=np.uint8), seq_off=seq_off,
                           l_seq=self.l_seq[sel].astype(np.int32), names=names, tags=tags_out)


def records_from_dicts(records, dtypes=None):
    """Build a RecordBatch from plain dict records (the reference's ``simple_data`` style): keys
    'pos', 'cigar' (string), 'seq', 'flag', 'mapq', optional 'tags'."""
    import re
    code = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
    opc = {c: i for i, c in enumerate("MIDNSHP=X")}
    cig, cig_off, seq, seq_off = [], [0], [], [0]
    dt = []
    for r in records:
        ops = [(int(n) << 4) | opc[o] for n, o in re.findall(r"(\d+)([MIDNSHP=X])", r["cigar"])]
        cig.extend(ops)
        cig_off.append(len(cig))
        s = r["seq"].upper()
        nib = [code.get(ch, 15) for ch in s] + ([0] if len(s) % 2 else [])
        seq.extend((nib[i] << 4) | nib[i + 1] for i in range(0, len(nib), 2))
        seq_off.append(len(seq))
        dt.append(list(dtypes).index(r["tags"]["DT"]) if dtypes is not None and len(dtypes) > 1 else 0)
    return RecordBatch(
        pos=np.array([r["pos"] for r in records], dtype=np.int32),
        flag=np.array([r.get("flag", 0) for r in records], dtype=np.uint16),
        mapq=np.array([r.get("mapq", 60) for r in records], dtype=np.uint8),
        dtype=np.array(dt, dtype=np.uint8), cigar=np.array(cig, dtype=np.uint32),
        cigar_off=np.array(cig_off, dtype=np.int64), seq=np.array(seq, dtype=np.uint8),
        seq_off=np.array(seq_off, dtype=np.int64), l_seq=np.array([len(r["seq"]) for r in records], dtype=np.int32),
        names=[r.get("query_name") for r in records], tags=[r.get("tags", {}) for r in records])
