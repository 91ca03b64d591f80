"""Oracle for count normalisation: CountsFeatureEncoder._post_process_pileup restated in numpy.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows medaka/features.py:871-935 (depth, minor columns inherit the depth of their
major column via searchsorted, optional sym_indels fill, normalise in
{'total','fwd_rev',None}, cast to float32) and medaka/features.py:647-687
(pileup_counts_norm_indices).  Constants from src/medaka_counts.h:19-22.
"""
from collections import defaultdict

import numpy as np

PLP_BASES = "acgtACGTdD"   # src/medaka_counts.h:19 ; lower case = reverse strand
FEATLEN = 10               # src/medaka_counts.h:20
FWD_DEL = 9                # src/medaka_counts.h:21
REV_DEL = 8                # src/medaka_counts.h:22


def pileup_counts_norm_indices(dtypes, num_qstrat=1):
    """medaka/features.py:647-687."""
    indices = defaultdict(list)
    for dti, dt in enumerate(dtypes):
        for qindex in range(num_qstrat):
            for base_i, code in enumerate(PLP_BASES):
                is_rev = code.islower()
                indices[dt, is_rev].append(
                    base_i + dti * num_qstrat * FEATLEN + qindex * FEATLEN)
    return dict(indices)


def post_process_pileup(counts, positions, normalise="total", dtypes=("",), sym_indels=False):
    """medaka/features.py:871-935.  Returns (features float32 [n,F], depth [n]).

    ``counts`` is modified in place when sym_indels is set, as the reference does.
    """
    feature_indices = pileup_counts_norm_indices(dtypes)
    minor_inds = np.where(positions["minor"] > 0)
    major_pos_at_minor_inds = positions["major"][minor_inds]
    major_ind_at_minor_inds = np.searchsorted(
        positions["major"], major_pos_at_minor_inds, side="left")

    depth = np.sum(counts, axis=1)
    depth[minor_inds] = depth[major_ind_at_minor_inds]

    if sym_indels:
        for (dt, is_rev), inds in feature_indices.items():
            dt_depth = np.sum(counts[:, inds], axis=1)
            featlen_index = REV_DEL if is_rev else FWD_DEL
            del_ind = [x for x in inds if x % FEATLEN == featlen_index][0]
            counts[minor_inds, del_ind] = \
                dt_depth[major_ind_at_minor_inds] - dt_depth[minor_inds]

    if normalise == "total":
        feature_array = counts / np.maximum(1, depth).reshape((-1, 1))
    elif normalise == "fwd_rev":
        feature_array = np.empty_like(counts, dtype=np.float32)
        for (dt, is_rev), inds in feature_indices.items():
            dt_depth = np.sum(counts[:, inds], axis=1)
            dt_depth[minor_inds] = dt_depth[major_ind_at_minor_inds]
            feature_array[:, inds] = \
                counts[:, inds] / np.maximum(1, dt_depth).reshape((-1, 1))
    else:
        feature_array = counts
    feature_array = feature_array.astype(np.float32)
    return feature_array, depth


def enforce_pileup_chunk_contiguity(pileups):
    """medaka/features.py:111-164: split on major gaps, re-join abutting chunks."""
    split_results = []
    for counts, positions in pileups:
        move = np.ediff1d(positions["major"])
        gaps = np.where(move > 1)[0] + 1
        if len(gaps) == 0:
            split_results.append((counts, positions))
        else:
            start = 0
            for i in gaps:
                split_results.append((counts[start:i], positions[start:i]))
                start = i
            split_results.append((counts[start:], positions[start:]))
    counts_buffer, positions_buffer, chunk_results, last = [], [], [], None
    for counts, positions in split_results:
        if len(positions) == 0:
            continue
        first = positions["major"][0]
        if len(counts_buffer) == 0 or first - last == 1:
            counts_buffer.append(counts)
            positions_buffer.append(positions)
            last = positions["major"][-1]
        else:
            chunk_results.append((np.concatenate(counts_buffer), np.concatenate(positions_buffer)))
            counts_buffer, positions_buffer = [counts], [positions]
            last = positions["major"][-1]
    if len(counts_buffer) != 0:
        chunk_results.append((np.concatenate(counts_buffer), np.concatenate(positions_buffer)))
    return chunk_results
