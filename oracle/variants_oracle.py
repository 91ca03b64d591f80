"""Oracle for variant decoding: numpy restatement of the reference.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows
  * medaka/labels.py:889-1014  HaploidLabelScheme.decode_variants: consensus with gaps -> reference with gaps ->
    variant columns -> runs -> per-run (ref, alt) strings and the log-likelihood-ratio quality
    sum(phred(1 - p_pred)) - sum(phred(1 - p_ref)), everything in the dtype of label_probs (float32 in production)
    with Python's left-to-right sum over the run;
  * src/medaka_rnn_variants.c:28-55 variant_columns (oracle/labels_oracle.py, also compiled from the reference's own
    source into oracle/_ref by oracle/Makefile);
  * medaka/vcf.py:338-415  Variant.trim / normalize (left-aligned, parsimonious ref/alt);
  * medaka/variant.py:30-119  join_samples: samples are re-cut so that no variant straddles a sample edge.
Pinned by tests/golden/variants.npz (made by tests/golden/make_variant_golden.py from the imported reference) and by
the reference's literal cases (medaka/test/test_labels.py:279-399).
"""
import numpy as np

from oracle import labels_oracle

SYMBOLS = '*ACGT'      # labels.py:342
ENC = {s: i for i, s in enumerate(SYMBOLS)}


def phred(err, cap=70.0):
    """labels.py:387-401, in the dtype of `err`."""
    err = np.clip(err, 10 ** (-cap / 10.0), 1)
    return np.minimum(-10 * np.log10(err), cap)


def pfmt(p, dp=3):
    """labels.py:404-416."""
    return '{:.{dp}f}'.format(round(p, dp), dp=dp)


def reference_with_gaps(positions, ref_seq):
    """labels.py:920-924: '*' on insertion columns, the draft base on major columns."""
    reference = np.full(len(positions), '*', dtype='|U1')
    reference[positions['minor'] == 0] = np.fromiter(
        ref_seq[positions['major'][0]:positions['major'][-1] + 1], dtype='|U1')
    return reference


def trim_variant(pos, ref, alts, reference):
    """vcf.py:338-402 with a reference (trim_end_and_align, then trim_start)."""
    seqs = [ref] + list(alts)
    changed = True
    while changed:
        changed = False
        if all(len(s) > 0 for s in seqs) and len(set(s[-1] for s in seqs)) == 1:
            seqs = [s[:-1] for s in seqs]
            changed = True
        if any(len(s) == 0 for s in seqs):
            if pos == 0:
                seqs = [s + reference[len(seqs[0])] for s in seqs]
                break
            pos -= 1
            seqs = [reference[pos] + s for s in seqs]
            changed = True
    min_len = min(len(s) for s in seqs)
    trim = 0
    for bases in zip(*seqs):
        same = len(set(bases)) == 1
        if not same or trim == min_len - 1:
            break
        trim += 1
    seqs = [s[trim:] for s in seqs]
    return pos + trim, seqs[0], seqs[1:]


def decode_variants(positions, label_probs, ref_seq, ambig_ref=False):
    """-> list of dict(pos, ref, alt, qual (float), qual_str, gq) like the Variant records of labels.py:889-989."""
    if positions['minor'][0] != 0:
        raise ValueError("The first position of a sample must not be an insertion.")
    probs = np.asarray(label_probs)
    mp = np.argmax(probs, -1)
    predicted = np.array(list(SYMBOLS), dtype='|U1')[mp]
    reference = reference_with_gaps(positions, ref_seq)
    is_variant = labels_oracle.variant_columns(
        positions['minor'], np.array([ord(c) for c in reference]), np.array([ord(c) for c in predicted]))
    out = []
    n = len(positions)
    i = 0
    while i < n:
        if not is_variant[i]:
            i += 1
            continue
        j = i
        while j < n and is_variant[j]:
            j += 1
        rstart, rend = i, j
        i = j
        ref_g = ''.join(reference[rstart:rend])
        pred_g = ''.join(predicted[rstart:rend])
        var_ref, var_pred = ref_g.replace('*', ''), pred_g.replace('*', '')
        if var_ref == var_pred:
            continue
        if not ambig_ref and not set(var_ref).issubset(set(SYMBOLS)):
            continue
        var_probs = probs[rstart:rend]
        ref_probs = np.array([var_probs[k, ENC[s if s != 'N' else '*']] for k, s in enumerate(ref_g)])
        pred_probs = np.array([var_probs[k, ENC[s]] for k, s in enumerate(pred_g)])
        ref_quals, pred_quals = phred(1.0 - ref_probs), phred(1.0 - pred_probs)
        qual = sum(pred_quals) - sum(ref_quals)
        var_pos = int(positions['major'][rstart])
        if positions['minor'][rstart] != 0:
            var_ref = ref_seq[var_pos] + var_ref
            var_pred = ref_seq[var_pos] + var_pred
        # Variant(...).normalize(reference=ref_seq): Variant.__init__ upper-cases ref, qual becomes float(pfmt(qual))
        pos2, ref2, alt2 = var_pos, var_ref.upper(), [var_pred]
        if not all(x == ref2 for x in alt2):
            pos2, ref2, alt2 = trim_variant(pos2, ref2, alt2, ref_seq)
        out.append(dict(pos=int(pos2), ref=ref2, alt=alt2[0], qual=float(pfmt(qual)), gq=pfmt(qual, 0),
                        run=(rstart, rend)))
    return out


def join_cuts(samples, ref_seq):
    """variant.py:30-119 on a list of (positions, label_probs, is_last_in_contig): returns, per yielded sample, the list
    of (input sample index, row slice) pieces it is concatenated from."""
    queue, out = [], []
    for idx, (pos, probs, is_last) in enumerate(samples):
        n = len(pos)
        if is_last:
            queue.append((idx, 0, n))
            out.append(queue)
            queue = []
            continue
        call = np.array(list(SYMBOLS), dtype='|U1')[np.argmax(probs, -1)]
        refg = np.array([ref_seq[p['major']] if p['minor'] == 0 else '*' for p in pos], dtype='|U1')
        is_diff = call != refg
        both_gap = np.logical_and(call == '*', refg == '*')
        if np.all(np.logical_or(is_diff, both_gap)):
            queue.append((idx, 0, n))
            continue
        major_inds = np.where(pos['minor'] == 0)
        major_pos = pos['major'][major_inds]
        is_diff = call[major_inds] != refg[major_inds]
        for offset, d in enumerate(is_diff[::-1]):
            if not d:
                break
        last_non_var_pos = major_pos[-1 - offset]
        cut = int(np.searchsorted(pos['major'], last_non_var_pos, side='left'))
        to_yield = queue
        if cut > 0:
            to_yield = to_yield + [(idx, 0, cut)]
        if to_yield:
            out.append(to_yield)
        queue = [(idx, cut, n)]
    if queue:
        raise ValueError('Reached end of generator without is_last_in_contig being True')
    return out
