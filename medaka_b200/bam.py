"""BAM access for the GPU pileup-counts featuriser.

The reference reads alignments through htslib (``bam_itr_querys`` / ``bam_mplp_auto``,
src/medaka_counts.c:233-251), which is not part of its tree and not available here; the
engine does the per-base work on the GPU instead, so the host only has to (1) inflate the
BGZF blocks of the requested region and (2) slice out, per alignment record, the fields the
featuriser consumes in BAM's own packed encodings: 32-bit CIGAR ops (``len << 4 | op``) and
4-bit sequence codes.  Both are done natively (libmedaka_b200, csrc/bam_io.cu: zlib thread
pool, .bai index, CG-tag long CIGARs); this module wraps the result in a ``RecordBatch`` of
flat arrays ready for ``mdk_pileup_counts`` and applies the tag-based read filters.
``bgzf_decompress`` (whole-file inflate in Python) is kept for tests and small tools.
"""
import collections
import concurrent.futures
import os
import struct
import zlib

import numpy as np

# flags the pileup never sees (src/medaka_bamiter.c:19): UNMAP | SECONDARY | QCFAIL | DUP | SUPPLEMENTARY
FILTER_FLAGS = 0x4 | 0x100 | 0x200 | 0x400 | 0x800
_CONSUMES_REF = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0], dtype=np.int64)    # M I D N S H P = X
_CONSUMES_QRY = np.array([1, 1, 0, 0, 1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0], dtype=np.int64)

RecordBatch = collections.namedtuple(
    "RecordBatch",
    ["pos", "flag", "mapq", "dtype", "cigar", "cigar_off", "seq", "seq_off", "l_seq", "names", "tags", "qual", "aux",
     "aux_off"])
# qual: base qualities, l_seq bytes per read back to back (offsets = cumsum(l_seq)); aux / aux_off: the raw optional
# fields (move table, haplotag).  Only the read-level featuriser asks for them.
RecordBatch.__new__.__defaults__ = (None, None, None)


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


def _parse_tags(raw, arrays=False):
    """Aux fields -> dict (the scalar / string types the read filters look at; B arrays only when ``arrays``: the
    move table 'mv' of the read-level featuriser)."""
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
            width = {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}[sub]
            if arrays:
                tags[tag] = list(struct.unpack_from("<%d%s" % (cnt, {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i",
                                                                  "I": "I", "f": "f"}[sub]), raw, i + 5))
            i += 5 + cnt * width
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
    """An open BAM file: region fetches go through the native reader in libmedaka_b200 (csrc/bam_io.cu: threaded BGZF
    inflate, BAI-indexed access, bounded memory), the counterpart of the reference's ``bam_fset`` + ``bam_itr_querys``
    (src/medaka_bamiter.c:52-63, src/medaka_counts.c:233).  Only the records of the requested region are ever inflated.
    """

    def __init__(self, path, threads=4, index=None):
        from medaka_b200 import libmedaka as _lm
        self._lm = _lm
        lib, ffi = _lm.load(), _lm.ffi
        self.path = path
        self.threads = int(threads)
        pb = ffi.new("mdk_bam **")
        _lm.check(lib.mdk_bam_open(os.fsencode(path), os.fsencode(index) if index else ffi.NULL, pb))
        self._h = pb[0]
        n = lib.mdk_bam_n_refs(self._h)
        self.references = [ffi.string(lib.mdk_bam_ref_name(self._h, i)).decode() for i in range(n)]
        self.lengths = [int(lib.mdk_bam_ref_len(self._h, i)) for i in range(n)]
        self.has_index = bool(lib.mdk_bam_has_index(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._lm.lib.mdk_bam_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()

    def get_regions(self):
        """(name, length) of every reference sequence - what get_bam_regions needs (medaka/common.py:762-790)."""
        return list(zip(self.references, self.lengths))

    def fetch(self, ref_name, start, end, dtypes=None, tag_name=None, tag_value=None, keep_missing=False,
              read_group=None, with_names=False, min_mapq=0, exclude_flags=FILTER_FLAGS, with_qual=False,
              with_tags=False, with_aux=False):
        """Records overlapping [start, end) on ref_name that pass the read filter of src/medaka_bamiter.c:17-45, in the
        reference's order: flag and mapping quality first (native, before anything is parsed), then the tag, read-group
        and datatype tests on the survivors' aux fields."""
        lm = self._lm
        lib, ffi = lm.lib, lm.ffi
        tid = self.references.index(ref_name)
        if start is None:
            start = 0
        if end is None:
            end = self.lengths[tid]
        pbatch = ffi.new("mdk_bam_batch **")
        lm.check(lib.mdk_bam_fetch(self._h, tid, int(start), int(end), int(exclude_flags), int(min_mapq), self.threads,
                                   pbatch))
        batch = pbatch[0]
        try:
            n = int(lib.mdk_bam_batch_size(batch))
            ptrs = {k: ffi.new(t) for k, t in (
                ("pos", "const int32_t **"), ("flag", "const uint16_t **"), ("mapq", "const uint8_t **"),
                ("l_seq", "const int32_t **"), ("cigar", "const uint32_t **"), ("cigar_off", "const int64_t **"),
                ("seq", "const uint8_t **"), ("seq_off", "const int64_t **"), ("aux", "const uint8_t **"),
                ("aux_off", "const int64_t **"), ("names", "const char **"), ("name_off", "const int64_t **"))}
            lm.check(lib.mdk_bam_batch_arrays(batch, *[ptrs[k] for k in (
                "pos", "flag", "mapq", "l_seq", "cigar", "cigar_off", "seq", "seq_off", "aux", "aux_off", "names",
                "name_off")]))

            def arr(key, dtype, count):
                if count == 0:
                    return np.zeros(0, dtype=dtype)
                return np.frombuffer(ffi.buffer(ptrs[key][0], count * np.dtype(dtype).itemsize), dtype=dtype).copy()

            cigar_off = arr("cigar_off", np.int64, n + 1)
            seq_off = arr("seq_off", np.int64, n + 1)
            aux_off = arr("aux_off", np.int64, n + 1)
            name_off = arr("name_off", np.int64, n + 1)
            pos, flag, mapq = arr("pos", np.int32, n), arr("flag", np.uint16, n), arr("mapq", np.uint8, n)
            l_seq = arr("l_seq", np.int32, n)
            cigar = arr("cigar", np.uint32, int(cigar_off[-1]))
            seq = arr("seq", np.uint8, int(seq_off[-1]))
            aux = bytes(ffi.buffer(ptrs["aux"][0], int(aux_off[-1]))) if n and aux_off[-1] else b""
            names_raw = bytes(ffi.buffer(ptrs["names"][0], int(name_off[-1]))) if n and name_off[-1] else b""
            qual = None
            if with_qual:
                pq, pqo = ffi.new("const uint8_t **"), ffi.new("const int64_t **")
                lm.check(lib.mdk_bam_batch_qual(batch, pq, pqo))
                nq = int(l_seq.sum())
                qual = (np.frombuffer(ffi.buffer(pq[0], nq), dtype=np.uint8).copy() if nq else np.zeros(0, dtype=np.uint8))
        finally:
            lib.mdk_bam_batch_free(batch)

        def name(i):
            return names_raw[int(name_off[i]):int(name_off[i + 1])].decode()

        aux_arr = np.frombuffer(aux, dtype=np.uint8).copy() if with_aux else None
        aux_off_out = aux_off if with_aux else None

        need_tags = with_tags or bool(tag_name) or read_group is not None or (dtypes is not None and len(dtypes) > 1)
        dtype = np.zeros(n, dtype=np.uint8)
        tags_out = None
        keep = np.ones(n, dtype=bool)
        if need_tags:
            tags_out = []
            for i in range(n):
                tg = _parse_tags(aux[int(aux_off[i]):int(aux_off[i + 1])], arrays=with_tags)
                tags_out.append(tg)
                keep[i] = _passes_tag_filters(tg, tag_name, tag_value, keep_missing, read_group)
                if keep[i] and dtypes is not None and len(dtypes) > 1:
                    if tg.get("DT") not in dtypes:
                        raise ValueError("Datatype not found for {}.".format(name(i)))
                    dtype[i] = list(dtypes).index(tg["DT"])
            tags_out = [t for t, k in zip(tags_out, keep) if k]
        if not keep.all():
            sel = np.flatnonzero(keep)
            n_cig = (cigar_off[1:] - cigar_off[:-1])[sel]
            new_coff = np.zeros(len(sel) + 1, dtype=np.int64)
            np.cumsum(n_cig, out=new_coff[1:])
            op_idx = np.arange(int(new_coff[-1])) - np.repeat(new_coff[:-1], n_cig) + np.repeat(cigar_off[:-1][sel], n_cig)
            cigar = cigar[op_idx] if len(op_idx) else np.zeros(0, dtype=np.uint32)
            n_sb = (seq_off[1:] - seq_off[:-1])[sel]
            new_soff = np.zeros(len(sel) + 1, dtype=np.int64)
            np.cumsum(n_sb, out=new_soff[1:])
            b_idx = np.arange(int(new_soff[-1])) - np.repeat(new_soff[:-1], n_sb) + np.repeat(seq_off[:-1][sel], n_sb)
            seq = seq[b_idx] if len(b_idx) else np.zeros(0, dtype=np.uint8)
            names = [name(i) for i in sel] if with_names else None
            if qual is not None:
                q_off = np.zeros(n + 1, dtype=np.int64)
                np.cumsum(l_seq, out=q_off[1:])
                n_q = l_seq[sel].astype(np.int64)
                new_qoff = np.zeros(len(sel) + 1, dtype=np.int64)
                np.cumsum(n_q, out=new_qoff[1:])
                q_idx = np.arange(int(new_qoff[-1])) - np.repeat(new_qoff[:-1], n_q) + np.repeat(q_off[:-1][sel], n_q)
                qual = qual[q_idx] if len(q_idx) else np.zeros(0, dtype=np.uint8)
            if aux_arr is not None:
                n_a = (aux_off[1:] - aux_off[:-1])[sel]
                new_aoff = np.zeros(len(sel) + 1, dtype=np.int64)
                np.cumsum(n_a, out=new_aoff[1:])
                a_idx = np.arange(int(new_aoff[-1])) - np.repeat(new_aoff[:-1], n_a) + np.repeat(aux_off[:-1][sel], n_a)
                aux_arr = aux_arr[a_idx] if len(a_idx) else np.zeros(0, dtype=np.uint8)
                aux_off_out = new_aoff
            pos, flag, mapq, l_seq, dtype = pos[sel], flag[sel], mapq[sel], l_seq[sel], dtype[sel]
            cigar_off, seq_off = new_coff, new_soff
        else:
            names = [name(i) for i in range(n)] if with_names else None
        return RecordBatch(pos=pos, flag=flag, mapq=mapq, dtype=dtype, cigar=np.ascontiguousarray(cigar, dtype=np.uint32),
                           cigar_off=cigar_off, seq=np.ascontiguousarray(seq, dtype=np.uint8), seq_off=seq_off,
                           l_seq=l_seq, names=names, tags=tags_out, qual=qual, aux=aux_arr, aux_off=aux_off_out)


def _encode_aux(tags):
    """dict -> raw BAM optional fields (ints as 'i', strings as 'Z', integer lists as a 'B' array of int8 when they
    fit, int32 otherwise) - what a reader would hand over for the same record."""
    out = bytearray()
    for key, v in (tags or {}).items():
        if isinstance(v, (int, np.integer)):
            out += key.encode() + b"i" + struct.pack("<i", int(v))
        elif isinstance(v, str):
            out += key.encode() + b"Z" + v.encode() + b"\x00"
        elif isinstance(v, (list, tuple, np.ndarray)) and all(float(x) == int(x) for x in v):
            small = all(-128 <= int(x) <= 127 for x in v)
            out += key.encode() + b"B" + (b"c" if small else b"i") + struct.pack("<I", len(v))
            out += struct.pack("<%d%s" % (len(v), "b" if small else "i"), *[int(x) for x in v])
    return bytes(out)


def records_from_dicts(records, dtypes=None):
    """Build a RecordBatch from plain dict records (the reference's ``simple_data`` style): keys
    'pos', 'cigar' (string), 'seq', 'flag', 'mapq', optional 'tags', 'qual' (list of ints; absent = 0xff)."""
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
        names=[r.get("query_name") for r in records], tags=[r.get("tags", {}) for r in records],
        qual=np.array([q for r in records for q in (r["qual"] if r.get("qual") is not None else [0xFF] * len(r["seq"]))],
                      dtype=np.uint8),
        aux=np.frombuffer(b"".join(_encode_aux(r.get("tags")) for r in records), dtype=np.uint8).copy(),
        aux_off=np.concatenate([[0], np.cumsum([len(_encode_aux(r.get("tags"))) for r in records])]).astype(np.int64))
