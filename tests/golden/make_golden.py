"""Generate golden vectors by running the REAL reference classes (build container only).

Run:  python tests/golden/make_golden.py        (needs /root/reference; writes tests/golden/*.npz)

The reference package cannot be imported as-is here (pysam, h5py, intervaltree,
libmedaka, ... are absent - SURVEY.md top table).  None of those are touched by the
functions on the hot path, so this script installs inert stand-ins for the missing
third-party modules, imports the reference from /root/reference UNMODIFIED, and
records the outputs of:

  * medaka.architectures.gru.GRUModel + medaka.models.TorchModel.predict_on_batch
  * medaka.features.CountsFeatureEncoder._post_process_pileup, pileup_counts_norm_indices
  * medaka.common.Sample.chunks, Region.split, sliding_window, grouper
  * medaka.labels.HaploidLabelScheme.decode_consensus (+ _phred)
  * medaka.torch_ext.Batch.collate

Inputs come from oracle/synth.py (RandomState, reproducible anywhere) or are the
reference's own literal test vectors (medaka/test/test_counts.py, test_labels.py).
The .npz files are committed; /root/reference is never read by the tests.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def install_stubs():
    class _LV:
        def __init__(self, v):
            self.v = tuple(int(x) for x in str(v).split('.') if x.isdigit())

        def __lt__(self, o):
            return self.v < o.v

        def __ge__(self, o):
            return self.v >= o.v

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    d = stub('distutils')
    d.version = stub('distutils.version', LooseVersion=_LV)
    stub('pkg_resources', resource_filename=lambda *a: '/nonexistent')

    class _FFI:
        def string(self, x):
            return x

        def buffer(self, x, n=None):
            return x[:n] if n else x

    class _Lib:  # constants of src/medaka_counts.h:19-22, src/medaka_read_matrix.h:37
        plp_bases = b'acgtACGTdD'
        featlen = 10
        fwd_del = 9
        rev_del = 8
        base_featlen = 4

    stub('libmedaka', ffi=_FFI(), lib=_Lib())

    class _Auto(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith('__'):
                raise AttributeError(k)
            return object

    class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        TOP = {'parasail', 'edlib', 'mappy', 'requests', 'h5py', 'toml', 'intervaltree', 'pysam',
               'ont_fast5_api', 'pyabpoa', 'wurlitzer', 'pyspoa', 'spoa', 'Bio', 'whatshap'}

        def find_spec(self, name, path, target=None):
            if name.split('.')[0] in self.TOP:
                return importlib.machinery.ModuleSpec(name, self, is_package=True)

        def create_module(self, spec):
            return _Auto(spec.name)

        def exec_module(self, m):
            m.__path__ = []

    sys.meta_path.append(_Finder())
    sys.path.insert(0, '/root/reference')
    import numpy as np
    if not hasattr(np, 'string_'):
        np.string_ = np.bytes_


def main():
    install_stubs()
    import numpy as np
    import torch
    import medaka.architectures.gru as ref_gru
    import medaka.common as ref_common
    import medaka.features as ref_features
    import medaka.labels as ref_labels
    import medaka.torch_ext as ref_torch_ext
    from oracle import synth

    torch.set_num_threads(8)
    meta = "medaka v%s, torch %s, numpy %s" % (
        __import__('medaka').__version__, torch.__version__, np.__version__)
    print(meta)

    # ---------------------------------------------------------------- forward pass
    cases = {}
    for name, seed, B, T, F, head_gain, rec_gain in [
            ("small", 0, 3, 64, 10, 8.0, 1.0),
            ("long", 1, 2, 1500, 10, 8.0, 1.0),
            ("hot", 2, 5, 300, 10, 24.0, 2.5),      # larger recurrent gain / sharper logits
            ("f20", 3, 2, 200, 20, 8.0, 1.0),       # two dtypes (config 5)
            ("b1", 4, 1, 777, 10, 8.0, 1.0),        # remainder path: B=1, odd T
            ("neartie", 5, 6, 400, 10, 8.0, 1.0)]:  # adversarial head: two classes within ~1e-5 of each other everywhere
        if name == "neartie":
            sd = synth.synth_state_dict_neartie(seed, num_features=F, head_gain=head_gain, rec_gain=rec_gain)
        else:
            sd = synth.synth_state_dict(seed, num_features=F, head_gain=head_gain, rec_gain=rec_gain)
        model = ref_gru.GRUModel(num_features=F)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        model.eval()
        feats = synth.synth_features(B, T, F, seed=100 + seed)

        class _Batch:
            counts_matrix = torch.from_numpy(feats)

        probs = model.predict_on_batch(_Batch())          # medaka/models.py:303-313
        assert probs.device.type == 'cpu' and probs.dtype == torch.float32
        model.normalise = False                            # gru.py:68-71 -> logits
        logits = model.predict_on_batch(_Batch())
        cases[name + "_probs"] = probs.numpy()
        cases[name + "_logits"] = logits.numpy()
        cases[name + "_args"] = np.array([seed, B, T, F, head_gain, rec_gain], dtype=np.float64)
        print("forward", name, probs.shape, float(probs.max()))
    np.savez_compressed(os.path.join(HERE, "gru_forward.npz"), meta=meta, **cases)

    # ---------------------------------------------------------------- normalisation
    out = {}
    # the reference's own golden (medaka/test/test_counts.py:298-311)
    simple_counts = np.array(
        [[2, 0, 0, 0, 2, 0, 0, 0, 0, 0],
         [0, 2, 0, 0, 0, 2, 0, 0, 0, 0],
         [2, 0, 0, 0, 2, 0, 0, 0, 0, 0],
         [0, 1, 0, 1, 0, 0, 0, 1, 0, 1],
         [1, 0, 0, 0, 0, 0, 0, 0, 0, 0],
         [0, 0, 2, 0, 0, 0, 2, 0, 0, 0],
         [2, 0, 0, 0, 2, 0, 0, 0, 0, 0],
         [0, 0, 0, 2, 0, 0, 0, 2, 0, 0],
         [0, 0, 2, 0, 0, 0, 2, 0, 0, 0]], dtype=np.uint64)
    simple_pos = np.array(
        [(0, 0), (1, 0), (2, 0), (3, 0), (3, 1), (4, 0), (5, 0), (6, 0), (7, 0)],
        dtype=[('major', '<i8'), ('minor', '<i8')])
    inputs = {
        "simple": (simple_counts, simple_pos, ('',)),
        "synth": synth.synth_counts(5000, seed=7) + (('',),),
        "synth_minor_start": synth.synth_counts(3000, seed=8, start_on_minor=True,
                                                start_major=1000) + (('',),),
        "synth2dt": synth.synth_counts(2000, seed=9, num_dtypes=2) + (('r9', 'r10'),),
        "deep": synth.synth_counts(1000, seed=10, mean_depth=20000, max_depth=100000) + (('',),),
    }
    region = ref_common.Region('ref', 0, 8)
    for name, (counts, pos, dtypes) in inputs.items():
        out[name + "_counts"] = counts
        out[name + "_major"] = pos['major']
        out[name + "_minor"] = pos['minor']
        for norm in ('total', 'fwd_rev', None):
            for sym in (False, True):
                enc = ref_features.CountsFeatureEncoder(
                    normalise=norm, dtypes=dtypes, sym_indels=sym)
                s = enc._post_process_pileup(counts.copy(), pos, region)
                key = "%s_%s_%d" % (name, norm, int(sym))
                out[key + "_features"] = s.features
                out[key + "_depth"] = np.asarray(s.depth)
    np.savez_compressed(os.path.join(HERE, "post_process.npz"), meta=meta, **out)
    idx = {}
    for dtypes, nq in ((('',), 1), (('1', '2'), 1), (('',), 2)):
        got = ref_features.pileup_counts_norm_indices(list(dtypes), num_qstrat=nq)
        for (dt, rev), v in got.items():
            idx["%s|%d|%s|%d" % (",".join(dtypes), nq, dt, int(rev))] = np.array(v)
    np.savez_compressed(os.path.join(HERE, "norm_indices.npz"), meta=meta, **idx)

    # ---------------------------------------------------------------- chunking / regions
    ch = {}
    for n, cl, ov in [(10, 4, 2), (100, 10, 3), (25000, 10000, 1000), (10000, 10000, 1000),
                      (10001, 10000, 1000), (19000, 10000, 1000), (19001, 10000, 1000),
                      (37, 5, 0), (12345, 1000, 200)]:
        pos = np.empty(n, dtype=[('major', '<i8'), ('minor', '<i8')])
        pos['major'] = np.arange(n)
        pos['minor'] = 0
        s = ref_common.Sample('c', np.arange(n), None, None, pos, None, None)
        starts = [(int(c.features[0]), int(c.features[-1]) + 1) for c in s.chunks(cl, ov)]
        ch["chunks_%d_%d_%d" % (n, cl, ov)] = np.array(starts, dtype=np.int64).reshape(-1, 2)
    for start, end, size, ov, fixed in [(0, 2500000, 1000000, 1000, False), (0, 100, 30, 5, True),
                                        (0, 100, 30, 5, False), (10, 50, 100, 0, True),
                                        (0, 8, 3, 0, False), (0, 300000, 100000, 0, False)]:
        r = ref_common.Region('c', start, end).split(size, overlap=ov, fixed_size=fixed)
        ch["split_%d_%d_%d_%d_%d" % (start, end, size, ov, int(fixed))] = np.array(
            [(x.start, x.end) for x in r], dtype=np.int64)
    ch["grouper_10_4"] = np.array([len(g) for g in ref_common.grouper(iter(range(10)), 4)])
    np.savez_compressed(os.path.join(HERE, "chunks.npz"), meta=meta, **ch)

    # ---------------------------------------------------------------- decode
    ls = ref_labels.HaploidLabelScheme()
    dec = {}

    class _S:
        pass

    s = _S()
    # medaka/test/test_labels.py:252-262
    s.label_probs = np.array([[0., 0.991, 0.009, 0., 0.], [0.1, 0., 0.9, 0., 0.],
                              [0.9, 0., 0.02, 0.04, 0.04], [0, 0, 0, 0, 1],
                              [0, 0.1, 0.1, 0.6, 0.2], [0, 0.01, 0.1, 0.88, 0.01]])
    seq, qual = ls.decode_consensus(s, with_qualities=True)
    assert (seq, qual) == ('ACTGG', '5+g$*')
    rs = np.random.RandomState(5)
    logits = rs.normal(0, 4, (20000, 5)).astype(np.float32)
    logits[::7] *= 4          # very confident columns -> q capped at 70
    p = torch.softmax(torch.from_numpy(logits), -1).numpy()
    p[100] = [0.2, 0.2, 0.2, 0.2, 0.2]      # exact tie: first max wins
    p[101] = [0.1, 0.4, 0.4, 0.05, 0.05]
    p[102] = [0, 0, 0, 0, 1]
    s.label_probs = p
    dec["probs"] = p
    seq, qual = ls.decode_consensus(s, with_qualities=True)
    seq_g, qual_g = ls.decode_consensus(s, with_gaps=True, with_qualities=True)
    dec["seq"] = np.frombuffer(seq.encode(), dtype=np.uint8)
    dec["qual"] = np.frombuffer(qual.encode(), dtype=np.uint8)
    dec["seq_gaps"] = np.frombuffer(seq_g.encode(), dtype=np.uint8)
    dec["qual_gaps"] = np.frombuffer(qual_g.encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "decode.npz"), meta=meta, **dec)

    # ---------------------------------------------------------------- collate
    feats = synth.synth_features(4, 50, 10, seed=77)
    samples = [ref_common.Sample('c', feats[i], None, None, None, None, None) for i in range(4)]
    b = ref_torch_ext.Batch.collate(samples)
    assert b.counts_matrix.dtype == torch.float32 and tuple(b.counts_matrix.shape) == (4, 50, 10)
    np.savez_compressed(os.path.join(HERE, "collate.npz"), meta=meta,
                        feats=feats, counts_matrix=b.counts_matrix.numpy())
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
