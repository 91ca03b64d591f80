"""Oracle for consensus stitching.

TEST INFRASTRUCTURE (see oracle/__init__.py) - never imported by the product path.

CPU restatement of what `medaka stitch` does with the samples of one region:

    medaka/stitch.py:33-85        _stitch_samples    decode every trimmed sample, cut contigs at the breaks
    medaka/common.py:495-557      trim_samples       remove the overlap between consecutive samples
    medaka/common.py:327-427      overlap_indices    where to cut (mid-point, or the junction heuristic)
    medaka/common.py:233-325      relative_position
    medaka/common.py:560-609      trim_samples_to_region
    medaka/common.py:612-644,85-98  filter_samples / depth_filter

A sample here is a plain dict {ref_name, positions (structured major/minor), label_probs [n,5], depth [n]}; every
stage materialises sliced copies, which is deliberately the opposite of the product's index-range planner
(medaka_b200/stitch.py).  Pinned by tests/golden/stitch.npz, recorded from the reference's own `_stitch_samples`
(tests/golden/make_stitch_golden.py).
"""
import numpy as np

from oracle import labels_oracle


class OverlapError(Exception):
    pass


def _first(s):
    p = s['positions'][0]
    return int(p['major']), int(p['minor'])


def _last(s):
    p = s['positions'][-1]
    return int(p['major']), int(p['minor'])


def _cut(s, lo, hi):
    return dict(ref_name=s['ref_name'], positions=s['positions'][lo:hi],
                label_probs=s['label_probs'][lo:hi], depth=s['depth'][lo:hi])


def relationship(s1, s2):
    """-> one of 'different_ref_name', '{forward,reverse}_{overlap,abutted,gapped}', 's2_within_s1', 's1_within_s2'."""
    if s1['ref_name'] != s2['ref_name']:
        return 'different_ref_name'
    a, b = s1, s2
    forward = (_first(s1), -len(s1['positions'])) <= (_first(s2), -len(s2['positions']))
    if not forward:
        a, b = s2, s1
    if _first(b) >= _first(a) and _last(b) <= _last(a):
        return 's2_within_s1' if forward else 's1_within_s2'
    end_maj, end_min = _last(a)
    st_maj, st_min = _first(b)
    if (st_maj == end_maj + 1 and st_min == 0) or (st_maj == end_maj and st_min == end_min + 1):
        kind = 'abutted'
    elif st_maj < end_maj or (st_maj == end_maj and st_min < end_min + 1):
        kind = 'overlap'
    elif st_maj > end_maj + 1 or (st_maj > end_maj and st_min > 0) or (st_maj == end_maj and st_min > end_min + 1):
        kind = 'gapped'
    else:
        raise RuntimeError('unclassifiable sample pair')
    return ('forward_' if forward else 'reverse_') + kind


def overlap_indices(s1, s2):
    """-> (end1, start2, heuristic): s1[:end1] + s2[start2:] is gapless and overlap-free (None = untouched)."""
    rel = relationship(s1, s2)
    if rel == 'forward_abutted':
        return None, None, False
    if rel != 'forward_overlap':
        raise OverlapError('cannot overlap samples related as ' + rel)
    p1, p2 = s1['positions'], s2['positions']
    first_in_1 = int(np.searchsorted(p1, p2[0]))
    past_in_2 = int(np.searchsorted(p2, p1[-1], side='right'))
    o1, o2 = p1[first_in_1:], p2[:past_in_2]
    if np.array_equal(o1['minor'], o2['minor']):
        n = len(o1)
        return first_in_1 + n // 2, past_in_2 - (n - n // 2), False
    # the two samples disagree on the columns of the overlap: look, outwards from the middle major position, for a
    # major position carrying equally many columns in both
    if len(np.unique(o1['major'])) > 3 and len(np.unique(o2['major'])) > 3:
        lo, hi = int(o1['major'][0]), int(o1['major'][-1])
        mid = lo + (hi - lo) // 2
        step = 1
        while not (mid + step > p1['major'].max() and mid - step < p2['major'].min()):
            for cand in (mid + step, mid - step):
                i1 = np.flatnonzero(p1['major'] == cand)
                i2 = np.flatnonzero(p2['major'] == cand)
                if len(i1) == len(i2):
                    return int(i1[0]), int(i2[0]), True
            step += 1
    raise OverlapError('no viable junction')


def trim_samples(samples):
    """list of samples -> list of (sample, is_last_in_contig, heuristic) with overlaps removed."""
    out = []
    it = iter(samples)
    try:
        cur = next(it)
    except StopIteration:
        return out
    cur_lo = None
    for nxt in list(it) + [None]:
        heuristic, last, nxt_lo = False, False, None
        if nxt is None:
            cur_hi, last = None, True
        else:
            rel = relationship(cur, nxt)
            if rel == 's2_within_s1':
                continue
            if rel == 'forward_gapped':
                cur_hi, last = None, True
            else:
                cur_hi, nxt_lo, heuristic = overlap_indices(cur, nxt)
        n = len(cur['positions'])
        lo, hi, _ = slice(cur_lo, cur_hi).indices(n)
        out.append((_cut(cur, lo, hi), last, heuristic))
        cur, cur_lo = nxt, nxt_lo
    return out


def trim_to_region(trimmed, start=None, end=None):
    out = []
    for s, last, heur in trimmed:
        maj = s['positions']['major']
        if start is not None:
            if maj[-1] < start:
                continue
            if maj[0] < start:
                probe = np.array([(start, 0)], dtype=s['positions'].dtype)[0]
                k = int(np.searchsorted(s['positions'], probe))
                s = _cut(s, k, len(maj))
                maj = s['positions']['major']
            if len(maj) == 0:
                continue
        if end is not None:
            if maj[0] >= end:
                break
            if maj[-1] >= end:
                s = _cut(s, 0, int(np.searchsorted(maj, end)))
            if len(s['positions']) == 0:
                continue
        out.append((s, last, heur))
    return out


def depth_filter(s, min_depth):
    """Runs of columns with depth >= min_depth, as separate samples."""
    ok = np.asarray(s['depth']) >= min_depth
    pieces, i, n = [], 0, len(ok)
    while i < n:
        j = i
        while j < n and ok[j] == ok[i]:
            j += 1
        if ok[i]:
            pieces.append(_cut(s, i, j))
        i = j
    return pieces


def stitch_samples(samples, start=None, end=None, min_depth=0):
    """-> list of ((ref_name, first major, last major), [sequence parts], [quality parts]) like _stitch_samples."""
    stream = trim_to_region(trim_samples(samples), start, end)
    if min_depth:
        kept = []
        for s, _, _ in stream:
            kept.extend(depth_filter(s, min_depth))
        stream = trim_samples(kept)
    contigs, seqs, quals, first = [], [], [], None
    s = None
    for s, last, _ in stream:
        if first is None:
            first = int(s['positions']['major'][0])
        seq, qual = labels_oracle.decode_consensus(s['label_probs'], with_qualities=True)
        seqs.append(seq)
        quals.append(qual)
        if last:
            contigs.append(((s['ref_name'], first, int(s['positions']['major'][-1])), seqs, quals))
            seqs, quals, first = [], [], None
    if seqs:
        contigs.append(((s['ref_name'], first, int(s['positions']['major'][-1])), seqs, quals))
    return contigs
