"""Record the reference's own stitching results (build container only; needs /root/reference).

Runs the UNMODIFIED `medaka.stitch._stitch_samples` (stitch.py:33-85, which drives Sample.trim_samples /
trim_samples_to_region / filter_samples and HaploidLabelScheme.decode_consensus) and `collapse_neighbours`
on the sample streams of oracle/synth.py::synth_stitch_stream, using the import stand-ins of make_golden.py, and
writes tests/golden/stitch.npz: per case the generator arguments (so the tests rebuild the inputs) and the contigs.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import make_golden  # noqa: E402

CASES = {
    # name: (stream kwargs, region (start, end), min_depth)
    "plain": (dict(seed=1), (None, None), 0),
    "region": (dict(seed=2), (1490, 3210), 0),
    "ragged": (dict(seed=3, ragged=(1, 4, 6)), (None, None), 0),
    "gapped": (dict(seed=4, drop=(3, 7), overlap=100), (None, None), 0),
    "nested": (dict(seed=5, nest=(2, 5)), (None, None), 0),
    "depth": (dict(seed=6, low_depth=((1, 150, 170), (4, 0, 30), (6, 380, 400))), (None, None), 10),
    "everything": (dict(seed=7, ragged=(2,), drop=(5,), nest=(3,), low_depth=((1, 150, 170), (8, 10, 50))),
                   (1200, 3900), 10),
    "no_insertions": (dict(seed=8, p_ins=0.0, n_major=1500, chunk_len=300, overlap=60), (1010, None), 0),
    "odd_overlap": (dict(seed=9, chunk_len=333, overlap=77), (None, 3500), 0),
    "single": (dict(seed=10, n_major=200, chunk_len=1000, overlap=100), (None, None), 0),
}


def main():
    make_golden.install_stubs()
    import numpy as np
    import medaka.common as ref_common
    import medaka.labels as ref_labels
    import medaka.stitch as ref_stitch
    from oracle import synth

    scheme = ref_labels.HaploidLabelScheme()
    out = {}
    for name, (kw, (start, end), min_depth) in CASES.items():
        stream = synth.synth_stitch_stream(**kw)
        samples = [ref_common.Sample(ref_name=s['ref_name'], features=None, labels=None, ref_seq=None,
                                     positions=s['positions'], label_probs=s['label_probs'], depth=s['depth'])
                   for s in stream]
        region = ref_common.Region('contig1', start, end)
        contigs = ref_stitch._stitch_samples(iter(samples), scheme, region, min_depth)
        joined = list(ref_stitch.collapse_neighbours(iter(
            [(c[0], list(c[1]), list(c[2])) for c in contigs])))
        rec = dict(kwargs=kw, start=start, end=end, min_depth=min_depth,
                   contigs=[[[c[0][0], int(c[0][1]), int(c[0][2])], ''.join(c[1]), ''.join(c[2]), [len(x) for x in c[1]]] for c in contigs],
                   collapsed=[[[c[0][0], int(c[0][1]), int(c[0][2])], ''.join(c[1]), ''.join(c[2])] for c in joined])
        out[name] = json.dumps(rec)
        print(name, [(c[0], len(''.join(c[1]))) for c in contigs])
    meta = "medaka v%s, numpy %s" % (__import__('medaka').__version__, np.__version__)
    np.savez_compressed(os.path.join(HERE, "stitch.npz"), meta=meta, **out)


if __name__ == "__main__":
    main()
