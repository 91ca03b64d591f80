"""Decode seam: per-position label decode on the GPU.

Mirrors ``HaploidLabelScheme.decode_consensus`` (medaka/labels.py:1053-1085) and ``_phred``
(:387-401).  The array work - argmax (first maximum wins), probability of the chosen class,
``uint8(min(70, -10*log10(clip(1-p, 1e-7, 1)))) + 33`` - runs in libmedaka_b200
(mdk_decode_consensus); gap removal and string building, which are O(n) byte shuffles on
the result, stay on the host.
"""
import numpy as np

from medaka_b200 import libmedaka as _lm


def decode_arrays(label_probs, device=0, with_qualities=True):
    """float32 probabilities [..., 5] -> (labels uint8 [...], quals uint8 [...] or None)."""
    lib = _lm.load()
    label_probs = label_probs.detach().cpu().numpy() if hasattr(label_probs, "detach") else np.asarray(label_probs)
    # numpy decodes in the array's own precision: float64 stays float64, everything else runs as float32
    f64 = label_probs.dtype == np.float64
    p = np.ascontiguousarray(label_probs, dtype=np.float64 if f64 else np.float32)
    if p.shape[-1] != 5:
        raise ValueError("expected label probabilities with 5 classes, got shape {}".format(p.shape))
    n = int(np.prod(p.shape[:-1]))
    labels = np.empty(p.shape[:-1], dtype=np.uint8)
    quals = np.empty(p.shape[:-1], dtype=np.uint8) if with_qualities else None
    ffi = _lm.ffi
    fn, ctype = (lib.mdk_decode_consensus_f64, "const double *") if f64 else (lib.mdk_decode_consensus, "const float *")
    _lm.check(fn(
        device, ffi.cast(ctype, ffi.from_buffer(p)), n,
        ffi.cast("uint8_t *", ffi.from_buffer(labels)),
        ffi.cast("uint8_t *", ffi.from_buffer(quals)) if with_qualities else ffi.NULL))
    return labels, quals


def variant_columns(minor, reference, prediction, device=0):
    """Which pileup columns belong to a variant run - ``HaploidLabelScheme._find_variants``
    (medaka/labels.py:869-887 -> libmedaka.lib.variant_columns, src/medaka_rnn_variants.c:28-55) on the GPU.

    :param minor: pileup minor indices; :param reference, prediction: per-column symbols (str, '|U1' / '|S1'
        arrays or uint8 label codes) including gaps.  :returns: bool array.
    """
    lib, ffi = _lm.load(), _lm.ffi

    def codes(x):
        a = np.asarray(list(x) if isinstance(x, str) else x)
        if a.dtype.kind == 'U':
            a = np.array([ord(c) for c in a.tolist()], dtype=np.uint32)
        elif a.dtype.kind == 'S':
            a = np.frombuffer(a.tobytes(), dtype=np.uint8)
        if a.size and int(a.max()) > 255:
            raise ValueError("symbols must fit one byte")
        return np.ascontiguousarray(a, dtype=np.uint8)

    mn = np.ascontiguousarray(minor, dtype=np.int64)
    r, p = codes(reference), codes(prediction)
    if not (len(mn) == len(r) == len(p)):
        raise ValueError("minor, reference and prediction must have the same length")
    out = np.zeros(len(mn), dtype=np.uint8)
    _lm.check(lib.mdk_variant_columns(device, ffi.cast("const int64_t *", ffi.from_buffer(mn)),
                                      ffi.cast("const uint8_t *", ffi.from_buffer(r)),
                                      ffi.cast("const uint8_t *", ffi.from_buffer(p)),
                                      ffi.cast("uint8_t *", ffi.from_buffer(out)), len(mn)))
    return out.astype(bool)


def decode_variant_arrays(label_probs, minor, ref_codes, device=0, want_quals=True):
    """The array half of ``decode_variants`` on the GPU (libmedaka_b200 ``mdk_decode_variants``).

    :param label_probs: float [n, 5]; :param minor: pileup minor indices [n]; :param ref_codes: uint8 [n], the draft
        with gaps as label codes (0..4 '*ACGT', 5 'N', 6 other).
    :returns: dict(pred uint8 [n], is_var bool [n], pred_q / ref_q float32 [n], run_start / run_len int64 [r],
        run_pred_q / run_ref_q float32 [r]).
    """
    lib, ffi = _lm.load(), _lm.ffi
    p = np.ascontiguousarray(label_probs.detach().cpu().numpy() if hasattr(label_probs, "detach") else label_probs,
                             dtype=np.float32)
    n = len(p)
    mn = np.ascontiguousarray(minor, dtype=np.int64)
    rc = np.ascontiguousarray(ref_codes, dtype=np.uint8)
    if p.ndim != 2 or p.shape[1] != 5 or len(mn) != n or len(rc) != n:
        raise ValueError("label_probs [n,5], minor [n] and ref_codes [n] expected")
    pred = np.empty(n, dtype=np.uint8)
    is_var = np.empty(n, dtype=np.uint8)
    pq = np.empty(n, dtype=np.float32) if want_quals else None
    rq = np.empty(n, dtype=np.float32) if want_quals else None
    max_runs = max(16, n // 8)
    n_runs = ffi.new("int64_t *")
    while True:
        rs, rl = np.empty(max_runs, dtype=np.int64), np.empty(max_runs, dtype=np.int64)
        rp, rr = np.empty(max_runs, dtype=np.float32), np.empty(max_runs, dtype=np.float32)
        code = lib.mdk_decode_variants(
            device, ffi.cast("const float *", ffi.from_buffer(p)), ffi.cast("const int64_t *", ffi.from_buffer(mn)),
            ffi.cast("const uint8_t *", ffi.from_buffer(rc)), n, ffi.cast("uint8_t *", ffi.from_buffer(pred)),
            ffi.cast("uint8_t *", ffi.from_buffer(is_var)),
            ffi.cast("float *", ffi.from_buffer(pq)) if want_quals else ffi.NULL,
            ffi.cast("float *", ffi.from_buffer(rq)) if want_quals else ffi.NULL, max_runs,
            ffi.cast("int64_t *", ffi.from_buffer(rs)), ffi.cast("int64_t *", ffi.from_buffer(rl)),
            ffi.cast("float *", ffi.from_buffer(rp)), ffi.cast("float *", ffi.from_buffer(rr)), n_runs)
        if code == lib.MDK_ERR_NOMEM and int(n_runs[0]) > max_runs:
            max_runs = int(n_runs[0])            # enlarge and retry, like enlarge_plp_data for the pileup
            continue
        _lm.check(code)
        break
    r = int(n_runs[0])
    return dict(pred=pred, is_var=is_var.astype(bool), pred_q=pq, ref_q=rq, run_start=rs[:r], run_len=rl[:r],
                run_pred_q=rp[:r], run_ref_q=rr[:r])


class HaploidLabelScheme(object):
    """The decode half of the reference's HaploidLabelScheme (labels.py:703-1085)."""

    symbols = '*ACGT'     # labels.py:342
    n_elements = 1

    def __init__(self, device=0):
        self.device = device
        self.verbose = False
        # draft symbol -> label code; 5 = 'N' (its own symbol when compared, the gap class when scored), 6 = other
        self._ref_table = np.full(256, 6, dtype=np.uint8)
        for i, c in enumerate(self.symbols):
            self._ref_table[ord(c)] = i
        self._ref_table[ord('N')] = 5

    # pickles as an attribute-less object, like the reference's label scheme (plain object pickling of a class whose
    # instances carry no state); the lookup table and the device are rebuilt on load
    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.__init__()
        self.__dict__.update(state)

    @staticmethod
    def _pfmt(p, dp=3):
        """labels.py:404-416."""
        if isinstance(p, np.ndarray):
            return np.char.mod("%.{}f".format(dp), p)
        return '{:.{dp}f}'.format(round(p, dp), dp=dp)

    def decode_labels(self, sample):
        """argmax label codes (uint8, gaps kept) of a sample - the integer form of decode_consensus(with_gaps=True)."""
        return decode_arrays(sample.label_probs, self.device, with_qualities=False)[0]

    def encode_reference(self, ref_seq, majors):
        """Label codes of the draft at the given major positions."""
        majors = np.asarray(majors, dtype=np.int64)
        if len(majors) == 0:
            return np.zeros(0, dtype=np.uint8)
        lo, hi = int(majors[0]), int(majors[-1]) + 1
        window = np.frombuffer(ref_seq[lo:hi].encode('ascii', 'replace'), dtype=np.uint8)
        return self._ref_table[window[majors - lo]]

    def decode_variants(self, sample, ref_seq, ambig_ref=False, return_all=False):
        """Convert network output in sample to variant records (medaka/labels.py:889-1014).

        The consensus with gaps, the variant columns, the per-column qualities and the per-run sums come from the GPU
        (``mdk_decode_variants``); the strings of the (few) variant runs, the ref == alt / ambiguous-draft filters and
        the VCF normalisation are done here.  Returns a list of ``medaka_b200.variant.Variant``.
        """
        from medaka_b200.variant import Variant
        pos = sample.positions
        if pos['minor'][0] != 0:
            raise ValueError("The first position of a sample must not be an insertion.")
        is_major = pos['minor'] == 0
        ref_codes = np.zeros(len(pos), dtype=np.uint8)            # '*' on insertion columns (labels.py:920)
        ref_codes[is_major] = self.encode_reference(ref_seq, pos['major'][is_major])
        d = decode_variant_arrays(sample.label_probs, pos['minor'], ref_codes, self.device)
        sym = np.frombuffer((self.symbols + 'N?').encode(), dtype=np.uint8)
        major0 = int(pos['major'][0])
        variants = []
        for rstart, rlen, spq, srq in zip(d['run_start'], d['run_len'], d['run_pred_q'], d['run_ref_q']):
            rstart, rend = int(rstart), int(rstart + rlen)
            codes = ref_codes[rstart:rend]
            if np.any(codes == 6):
                # spell the run's draft from the sequence itself (any IUPAC symbol)
                ref_g = ''.join(ref_seq[int(m)] if mi == 0 else '*' for m, mi in zip(pos['major'][rstart:rend],
                                                                                     pos['minor'][rstart:rend]))
            else:
                ref_g = sym[codes].tobytes().decode()
            pred_g = sym[d['pred'][rstart:rend]].tobytes().decode()
            var_ref, var_pred = ref_g.replace('*', ''), pred_g.replace('*', '')
            if var_ref == var_pred:          # deletion followed by insertion of the same base (labels.py:942-944)
                continue
            if not set(var_ref).issubset(set(self.symbols)):
                if not ambig_ref:
                    continue
                if set(var_ref) - set(self.symbols) - {'N'}:
                    raise KeyError("draft symbol outside '*ACGTN' in a variant run at {}:{}".format(
                        sample.ref_name, int(pos['major'][rstart])))
            qual = np.float32(spq) - np.float32(srq)          # log likelihood ratio (labels.py:974-975), float32
            info = {}
            if self.verbose:
                info = {'ref_seq': ref_g, 'pred_seq': pred_g,
                        'ref_qs': ','.join(self._pfmt(float(q)) for q in d['ref_q'][rstart:rend]),
                        'pred_qs': ','.join(self._pfmt(float(q)) for q in d['pred_q'][rstart:rend]),
                        'ref_q': self._pfmt(float(srq)), 'pred_q': self._pfmt(float(spq)), 'n_cols': rend - rstart}
            genotype = {'GT': '1', 'GQ': self._pfmt(float(qual), 0)}
            var_pos = int(pos['major'][rstart])
            if pos['minor'][rstart] != 0:    # variant starts on an insertion column: prepend the draft base
                var_ref = ref_seq[var_pos] + var_ref
                var_pred = ref_seq[var_pos] + var_pred
            v = Variant(sample.ref_name, var_pos, var_ref, alt=var_pred, filt='PASS', info=info,
                        qual=self._pfmt(float(qual)), genotype_data=genotype)
            variants.append(v.normalize(reference=ref_seq))
        if return_all:
            # one record per reference position (labels.py:991-1012)
            quals = d['ref_q'][is_major]
            qf, qi = np.char.mod("%.3f", quals), np.char.mod("%d", np.rint(quals))
            bases = [ref_seq[int(m)] for m in pos['major'][is_major]]
            for p_, base, f_, i_ in zip(pos['major'][is_major], bases, qf, qi):
                variants.append(Variant(sample.ref_name, int(p_), base, alt='.', filt='.', info={}, qual=str(f_),
                                        genotype_data={'GT': '0', 'GQ': str(i_)}))
            variants.sort(key=lambda x: x.pos)
        del major0
        return variants

    @property
    def num_classes(self):
        return len(self.symbols)

    @staticmethod
    def _phred(err, cap=70.0):
        """Host restatement kept for API compatibility (labels.py:387-401); not used on the hot path."""
        err = np.clip(err, 10 ** (-cap / 10.0), 1)
        return np.minimum(-10 * np.log10(err), cap)

    @staticmethod
    def _find_variants(minor, reference, prediction):
        """labels.py:869-887."""
        return variant_columns(minor, reference, prediction)

    def decode_consensus(self, sample, with_gaps=False, dtype=None, with_qualities=False):
        """Convert network output to consensus sequence by argmax decoding.

        :param sample: object with a ``label_probs`` array [n, 5].
        :param with_gaps: include gap ("*") characters in output.
        :returns: str, consensus sequence, optionally: qualities
        """
        mp, quals = decode_arrays(sample.label_probs, self.device, with_qualities=with_qualities)
        if not with_gaps:
            keep = mp != self.symbols.index('*')
            mp = mp[keep]
            if with_qualities:
                quals = quals[keep]
        if dtype is None:
            table = np.frombuffer(self.symbols.encode(), dtype=np.uint8)
            seq = table[mp].tobytes().decode()
        else:
            seq = np.fromiter(self.symbols, dtype=dtype)[mp]
        if with_qualities:
            return seq, quals.tobytes().decode()
        return seq
