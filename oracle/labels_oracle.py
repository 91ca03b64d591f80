"""Oracle for the per-position label decode: HaploidLabelScheme.decode_consensus in numpy.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows medaka/labels.py:1053-1085 (argmax, first max wins; drop '*' unless
with_gaps; symbols '*ACGT' labels.py:342; qualities) and medaka/labels.py:387-401
(_phred: clip err to [1e-7, 1], -10*log10, cap 70; then astype('u1') truncation and
+33 at labels.py:1082-1083).
"""
import numpy as np

SYMBOLS = "*ACGT"   # medaka/labels.py:342


def phred(err, cap=70.0):
    err = np.clip(err, 10 ** (-cap / 10.0), 1)
    q = -10 * np.log10(err)
    return np.minimum(q, cap)


def decode_arrays(label_probs):
    """Per-position arrays before gap removal: (labels u8 [n], qual u8 [n] = phred+33)."""
    mp = np.argmax(label_probs, -1)
    probs = np.take_along_axis(label_probs, np.expand_dims(mp, -1), -1).squeeze(-1)
    qual = phred(1 - probs).astype("u1") + 33
    return mp.astype(np.uint8), qual.astype(np.uint8)


def decode_consensus(label_probs, with_gaps=False, with_qualities=False):
    mp = np.argmax(label_probs, -1)
    if with_qualities:
        probs = np.take_along_axis(label_probs, np.expand_dims(mp, -1), -1).squeeze(-1)
    if not with_gaps:
        mask = mp != SYMBOLS.index("*")
        mp = mp[mask]
    decode = np.array([ord(x) for x in SYMBOLS], dtype="u1")
    seq = decode[mp].tobytes().decode()
    if with_qualities:
        if not with_gaps:
            probs = probs[mask]
        qual = (phred(1 - probs).astype("u1") + 33).tobytes().decode()
        return seq, qual
    return seq


def variant_columns(minor, reference, prediction):
    """src/medaka_rnn_variants.c:28-55 restated: sequential walk over the pileup columns."""
    n = len(minor)
    out = np.zeros(n, dtype=bool)
    if n == 0:
        return out
    is_var = reference[0] != prediction[0]      # assume start on major
    insert_length = 0
    out[0] = is_var
    for i in range(1, n):
        if minor[i] == 0:
            if is_var:
                out[i - insert_length:i] = True
            is_var = reference[i] != prediction[i]
            out[i] = is_var
            insert_length = 0
        else:
            insert_length += 1
            is_var = is_var or reference[i] != prediction[i]
    if is_var:
        out[n - insert_length:n] = True
    return out
