"""Pin the pileup oracle and the BAM reader on the reference's REAL test BAM (build container only).

Run:  python tests/golden/make_pileup_golden.py     (needs /root/reference; writes tests/golden/pileup_real.npz)

1. Reads /root/reference/medaka/test/data/test_reads.bam with medaka_b200.bam (no htslib) and runs
   oracle/pileup_oracle.py over utg000001l:50000-100000.  The reference's own regression test
   (medaka/test/test_counts.py:28-45, produced by its htslib-based calculate_pileup) expects: 86 294 columns,
   first position (50000, 0), last (99999, 1), first row [0,22,0,0,0,15,0,0,0,0], mean depth 18.696468.
   This script ASSERTS all of them, i.e. the restated bam_mplp_auto / resolve_cigar semantics are pinned on
   real data, not only on the 4-read mock.
2. Stores a small slice (the records overlapping utg000001l:50000-50250, BAM-packed) with the oracle's counts
   for that sub-region as the fixture the CPU and GPU tests replay.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from medaka_b200 import bam  # noqa: E402
from oracle import pileup_oracle  # noqa: E402

BAM = "/root/reference/medaka/test/data/test_reads.bam"


def main():
    bf = bam.BamFile(BAM)
    assert bf.references == ["utg000001l"]
    rb = bf.fetch("utg000001l", 50000, 100000)
    counts, pos = pileup_oracle.pileup_counts_from_batch(rb, 50000, 100000)
    assert counts.shape == (86294, 10)
    assert tuple(pos[0]) == (50000, 0) and tuple(pos[-1]) == (99999, 1)
    assert counts[0].tolist() == [0, 22, 0, 0, 0, 15, 0, 0, 0, 0]
    np.testing.assert_almost_equal(counts.sum(axis=1).mean(), 18.696468, decimal=6)
    print("real-BAM regression numbers reproduced: width 86294, mean depth %.6f" % counts.sum(axis=1).mean())

    start, end = 50000, 50250
    rb = bf.fetch("utg000001l", start, end)
    c, p = pileup_oracle.pileup_counts_from_batch(rb, start, end)
    np.savez_compressed(
        os.path.join(HERE, "pileup_real.npz"),
        meta="test_reads.bam utg000001l:%d-%d, %d records; full-region check 86294 cols / 18.696468 passed" % (
            start, end, len(rb.pos)),
        start=start, end=end, pos=rb.pos, flag=rb.flag, mapq=rb.mapq, dtype=rb.dtype, cigar=rb.cigar,
        cigar_off=rb.cigar_off, seq=rb.seq, seq_off=rb.seq_off, l_seq=rb.l_seq,
        counts=c, major=p["major"], minor=p["minor"])
    print("fixture:", len(rb.pos), "records,", c.shape, "columns")


if __name__ == "__main__":
    main()
