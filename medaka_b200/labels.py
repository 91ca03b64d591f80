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


class HaploidLabelScheme(object):
    """The decode half of the reference's HaploidLabelScheme (labels.py:703-1085)."""

    symbols = '*ACGT'     # labels.py:342
    n_elements = 1

    def __init__(self, device=0):
        self.device = device

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
