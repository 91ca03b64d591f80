"""Record the reference's own variant decoding (build container only; needs /root/reference and `make -C oracle`).

Runs the UNMODIFIED `medaka.labels.HaploidLabelScheme.decode_variants` (labels.py:889-1014; its Variant.normalize is
medaka/vcf.py:338-415) and `medaka.variant.join_samples` (variant.py:30-119, fed by Sample.trim_samples) with the
import stand-ins of make_golden.py.  `libmedaka.lib.variant_columns` - the one C function on this path - is the
reference's own src/medaka_rnn_variants.c compiled by oracle/Makefile into oracle/_ref/ and called through ctypes.
Writes tests/golden/variants.npz: per case the generator arguments (the tests rebuild the inputs from oracle/synth.py)
and the variant records.
"""
import ctypes
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import make_golden  # noqa: E402

DECODE_CASES = {
    "plain": dict(seed=1, n_major=3000),
    "dense": dict(seed=2, n_major=2000, p_mut=0.12, p_extend=0.5, p_ins_call=0.35),
    "no_ins_cols": dict(seed=3, n_major=1500, p_ins_col=0.0),
    "ambiguous": dict(seed=4, n_major=2500, n_frac=0.05),
    "at_origin": dict(seed=5, n_major=800, first_major=0, p_mut=0.08),
    "quiet": dict(seed=6, n_major=1000, p_mut=0.0, p_ins_call=0.0),
}
JOIN_CASES = {
    "join_plain": (dict(seed=11, n_major=4000, p_mut=0.05), 500, 100),
    "join_dense": (dict(seed=12, n_major=3000, p_mut=0.2, p_extend=0.6, p_ins_call=0.4), 300, 60),
    "join_short": (dict(seed=13, n_major=1200, p_mut=0.1), 150, 30),
}


def install_variant_columns():
    """libmedaka.ffi / libmedaka.lib as labels._find_variants uses them (labels.py:869-887), backed by the reference's C."""
    so = os.path.join(ROOT, "oracle", "_ref", "libmedaka_rnn_variants.so")
    cdll = ctypes.CDLL(so)
    cdll.variant_columns.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_size_t]
    cdll.variant_columns.restype = None
    lm = sys.modules['libmedaka']

    class FFI(type(lm.ffi)):
        def cast(self, ctype, value):
            return int(value)

    lm.ffi = FFI()
    lm.lib.variant_columns = staticmethod(
        lambda a, b, c, d, n: cdll.variant_columns(a, b, c, d, n))


def records(variants):
    return [dict(pos=int(v.pos), ref=v.ref, alt=list(v.alt), qual=float(v.qual), gt=v.genotype_data['GT'],
                 gq=v.genotype_data['GQ']) for v in variants]


def chunk_stream(d, chunk_len, overlap):
    import numpy as np
    n = len(d['positions'])
    step = chunk_len - overlap
    ranges = [(lo, lo + chunk_len) for lo in range(0, n - chunk_len + 1, step)]
    if not ranges or ranges[-1][1] < n:
        ranges.append((max(0, n - chunk_len), n))
    return ranges


def main():
    make_golden.install_stubs()
    install_variant_columns()
    import numpy as np
    import medaka.common as ref_common
    import medaka.labels as ref_labels
    import medaka.variant as ref_variant
    from oracle import synth

    ls = ref_labels.HaploidLabelScheme()
    out = {}
    for name, kw in DECODE_CASES.items():
        d = synth.synth_variant_pileup(**kw)
        s = ref_common.Sample(ref_name=d['ref_name'], features=None, labels=None, ref_seq=None,
                              positions=d['positions'], label_probs=d['label_probs'], depth=None)
        rec = dict(kwargs=kw)
        for ambig in (False, True):
            rec["ambig%d" % ambig] = records(ls.decode_variants(s, d['ref_seq'], ambig_ref=ambig))
        allv = ls.decode_variants(s, d['ref_seq'], return_all=True)
        rec["return_all_n"] = len(allv)
        rec["return_all_head"] = [dict(pos=int(v.pos), ref=v.ref, alt=list(v.alt), qual=float(v.qual) if v.qual != '.' else None,
                                       gt=v.genotype_data['GT'], gq=v.genotype_data['GQ']) for v in allv[:400]]
        out[name] = json.dumps(rec)
        print(name, len(rec["ambig0"]), len(rec["ambig1"]), rec["return_all_n"])
    for name, (kw, chunk_len, overlap) in JOIN_CASES.items():
        d = synth.synth_variant_pileup(**kw)
        ranges = chunk_stream(d, chunk_len, overlap)
        samples = [ref_common.Sample(ref_name=d['ref_name'], features=None, labels=None, ref_seq=None,
                                     positions=d['positions'][a:b], label_probs=d['label_probs'][a:b], depth=None)
                   for a, b in ranges]
        trimmed = ref_common.Sample.trim_samples(iter(samples))
        joined = list(ref_variant.join_samples(trimmed, d['ref_seq'], ls))
        rec = dict(kwargs=kw, chunk_len=chunk_len, overlap=overlap,
                   joined=[[s.name, int(s.size)] for s in joined], variants=[])
        for s in joined:
            rec["variants"].extend(records(ls.decode_variants(s, d['ref_seq'])))
        whole = ref_common.Sample(ref_name=d['ref_name'], features=None, labels=None, ref_seq=None,
                                  positions=d['positions'], label_probs=d['label_probs'], depth=None)
        rec["same_as_whole"] = rec["variants"] == records(ls.decode_variants(whole, d['ref_seq']))
        out[name] = json.dumps(rec)
        print(name, len(joined), len(rec["variants"]), rec["same_as_whole"])
    meta = "medaka v%s, numpy %s" % (__import__('medaka').__version__, np.__version__)
    np.savez_compressed(os.path.join(HERE, "variants.npz"), meta=meta, **out)


if __name__ == "__main__":
    main()
