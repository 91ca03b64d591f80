"""CPU tests of the host-side mirror (no GPU): containers, window scheduling, the DataLoader
contract of medaka/test/test_dataloader.py, output store round trips, multi-rank sharding."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from medaka_b200 import common, datastore, prediction, torch_ext

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _positions(n, start=0):
    pos = np.empty(n, dtype=[("major", "<i8"), ("minor", "<i8")])
    pos["major"] = np.arange(start, start + n)
    pos["minor"] = 0
    return pos


class FakeEncoder(object):
    """Stands in for CountsFeatureEncoder on hosts without a GPU: one fully covered contig,
    no insertions (the reference's test uses a single 5 kb read for the same reason)."""

    feature_vector_length = 10

    def bam_to_sample(self, bam, region):
        n = region.end - region.start
        feats = np.zeros((n, 10), dtype=np.float32)
        return [common.Sample(region.ref_name, feats, None, None, _positions(n, region.start), None,
                              np.ones(n, dtype=np.int64))]


def _run_loader(batch_size, chunk_len, chunk_overlap, regions=None):
    if regions is None:
        regions = [common.Region("ref", 0, 5000)]
    loader = prediction.DataLoader(
        None, regions, batch_size, batch_cache_size=4, bam_workers=4, feature_encoder=FakeEncoder(),
        chunk_len=chunk_len, chunk_overlap=chunk_overlap, enable_chunking=True)
    batches = [b for _, b in loader]
    return len(batches), sum(len(b.counts_matrix) for b in batches), len(loader.remainders)


# (batch_size, chunk_len, overlap, regions) -> (batches, samples, remainders): medaka/test/test_dataloader.py:49-89
FULL = common.Region("ref", 0, 5000)
DATALOADER_CASES = [
    ((200, 5000, 100, None), (1, 1, 0)),
    ((200, 10000, 100, None), (0, 0, 1)),
    ((200, 10000, 100, [FULL] * 5), (0, 0, 5)),
    ((200, 2500, 0, None), (1, 2, 0)),
    ((200, 2500, 100, None), (1, 3, 0)),
    ((2, 2500, 100, None), (2, 3, 0)),
    ((1, 5, 0, None), (1000, 1000, 0)),
    ((5000, 1, 0, None), (1, 5000, 0)),
    ((4999, 1, 0, None), (2, 5000, 0)),
    ((2, 1300, 0, [FULL] * 5 + [common.Region.from_string("ref:0-1000")] * 7), (10, 20, 7)),
    ((19, 250, 0, [FULL] * 100 + [common.Region.from_string("ref:0-100")] * 1000), (106, 2000, 1000)),
]


@pytest.mark.parametrize("args,expected", DATALOADER_CASES)
def test_dataloader_counts(args, expected):
    assert _run_loader(*args) == expected


def test_sample_basics_and_chunks():
    n = 25
    s = common.Sample("ctg", np.arange(n, dtype=np.float32).reshape(n, 1), None, None, _positions(n, 100),
                      None, np.arange(n))
    assert s.name == "ctg:100.0-124.0" and s.size == n and s.span == 24 and not s.is_empty
    assert common.Sample.decode_sample_name(s.name) == {"ref_name": "ctg", "start": "100.0", "end": "124.0"}
    chunks = list(s.chunks(chunk_len=10, overlap=3))
    assert [c.first_pos[0] for c in chunks] == [100, 107, 114, 115]       # last window right-aligned
    assert all(c.size == 10 for c in chunks)
    assert s.amend(depth=None).depth is None
    with pytest.raises(KeyError):
        s.amend(nonsense=1)
    assert s.slice(slice(2, 5)) == s.slice(slice(2, 5))
    assert s.slice(slice(2, 5)) != s.slice(slice(3, 6))


def test_relative_position():
    R = common.Relationship
    mk = lambda a, b, name="c": common.Sample(name, None, None, None, _positions(b - a, a), None, None)  # noqa
    assert common.Sample.relative_position(mk(0, 10), mk(10, 20)) is R.forward_abutted
    assert common.Sample.relative_position(mk(10, 20), mk(0, 10)) is R.reverse_abutted
    assert common.Sample.relative_position(mk(0, 10), mk(5, 20)) is R.forward_overlap
    assert common.Sample.relative_position(mk(5, 20), mk(0, 10)) is R.reverse_overlap
    assert common.Sample.relative_position(mk(0, 10), mk(12, 20)) is R.forward_gapped
    assert common.Sample.relative_position(mk(12, 20), mk(0, 10)) is R.reverse_gapped
    assert common.Sample.relative_position(mk(0, 20), mk(5, 10)) is R.s2_within_s1
    assert common.Sample.relative_position(mk(5, 10), mk(0, 20)) is R.s1_within_s2
    assert common.Sample.relative_position(mk(0, 10), mk(0, 10, "d")) is R.different_ref_name


def test_region_api(golden_dir):
    assert common.Region.from_string("Ecoli") == common.Region("Ecoli", None, None)
    assert common.Region.from_string("Ecoli:1000-2000") == common.Region("Ecoli", 1000, 2000)
    assert common.Region.from_string("Ecoli:-1000") == common.Region("Ecoli", 0, 1000)
    assert common.Region.from_string("A:B:c:500-") == common.Region("A:B:c", 500, None)
    assert str(common.Region("c", 5, 9)) == "c:5-9" and common.Region("c", 5, 9).size == 4
    g = np.load(os.path.join(golden_dir, "chunks.npz"))
    for key in g.files:
        parts = key.split("_")
        if parts[0] == "split":
            start, end, size, ov, fixed = map(int, parts[1:])
            got = [(r.start, r.end) for r in common.Region("c", start, end).split(size, overlap=ov, fixed_size=bool(fixed))]
            assert got == [tuple(x) for x in g[key].tolist()], key
    assert common.Region("a", 0, 10).overlaps(common.Region("a", 5, 15))
    assert not common.Region("a", 0, 10).overlaps(common.Region("b", 5, 15))


def test_grouper_and_rle():
    assert [len(x) for x in common.grouper(range(10), 4)] == [4, 4, 2]
    r = common.rle(np.array([1, 1, 2, 2, 2, 3]))
    assert r["length"].tolist() == [2, 3, 1] and r["start"].tolist() == [0, 2, 5] and r["value"].tolist() == [1, 2, 3]


def test_collate_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "collate.npz"))
    samples = [common.Sample("c", g["feats"][i], None, None, None, None, None) for i in range(4)]
    b = torch_ext.Batch.collate(samples)
    assert np.array_equal(b.counts_matrix.numpy(), g["counts_matrix"]) and b.features is b.counts_matrix
    out = np.empty((4, 50, 10), dtype=np.float32)
    b2 = torch_ext.Batch.collate(samples, out=out)
    assert b2.counts_matrix.data_ptr() == out.ctypes.data            # filled in place (pinned staging)
    bad = samples[:3] + [common.Sample("c", g["feats"][0][:10], None, None, None, None, None)]
    with pytest.raises(RuntimeError):
        torch_ext.Batch.collate(bad)


def test_triage_and_shard_regions():
    regs = [common.Region("a", 0, 2500000), common.Region("b", 0, 5000), common.Region("c", 0, 400000)]
    long, rem = prediction.triage_regions(regs, chunk_len=10000, bam_chunk=1000000, chunk_ovlp=1000)
    assert [r.ref_name for r in rem] == ["b"]
    assert [(r.start, r.end) for r in long if r.ref_name == "a"] == [(0, 1000000), (999000, 1999000), (1998000, 2500000)]
    shards = prediction.shard_regions(long, 2)
    assert sorted(r for s in shards for r in s) == sorted(long)
    loads = [sum(r.size for r in s) for s in shards]
    assert max(loads) - min(loads) <= max(r.size for r in long)


def test_datastore_roundtrip_and_resume():
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "probs.npzstore")
        n = 30
        s = common.Sample("ctg", None, None, None, _positions(n, 7), np.random.rand(n, 5).astype(np.float32),
                          np.arange(n))
        with datastore.DataStore(path, "a") as ds:
            ds.set_meta({"k": 1}, "model_function")
            ds.write_sample(s)
            ds.write_sample(s)                 # second write is a no-op (registry)
            assert ds.n_samples == 1
        with datastore.DataStore(path, "r") as ds:
            assert ds.sample_registry == {s.name}
            assert ds.get_meta("model_function") == {"k": 1}
            got = ds.load_sample(s.name)
            assert got.ref_name == "ctg" and np.array_equal(got.label_probs, s.label_probs)
            assert np.array_equal(got.positions, s.positions)
        with datastore.DataStore(path, "a") as ds:      # resume: already-present samples are skipped
            ds.write_sample(s.amend(label_probs=np.zeros((n, 5), dtype=np.float32)))
        with datastore.DataStore(path, "r") as ds:
            assert np.array_equal(ds.load_sample(s.name).label_probs, s.label_probs)


class StubModel(object):
    """The model seam without a GPU: records batch shapes, returns uniform probabilities; has the engine's async
    interface (predict_async / .result()) and its one-wave batch size."""

    def __init__(self, preferred, depth=3):
        self.preferred, self.shapes, self.depth = preferred, [], depth
        self.in_flight, self.max_in_flight, self.reserved = 0, 0, []

    def preferred_batch_size(self):
        return self.preferred

    def lookahead(self, batch_size, window_len=None):
        return self.depth

    def reserve(self, windows, window_len):
        self.reserved.append((windows, window_len))

    def predict_async(self, batch, slots=2):
        shape = tuple(batch.counts_matrix.shape)
        self.shapes.append(shape)
        assert slots > self.depth            # one staging slot per call in flight, plus the one being queued
        self.in_flight += 1
        self.max_in_flight = max(self.max_in_flight, self.in_flight)
        model = self

        class Handle(object):
            def result(_self):
                model.in_flight -= 1
                _self.labels = np.full(shape[:2], 3, dtype=np.uint8)
                return np.full(shape[:2] + (5,), 0.2, dtype=np.float32)
        return Handle()


def test_run_prediction_auto_batch_size_and_lookahead():
    """run_prediction (medaka/prediction.py:14-81) with batch_size="auto": batches take the engine's one-wave size, every
    window is written once, results come back in order through the look-ahead queue (at most `depth` batches in flight),
    and the engine's argmax labels are stored next to label_probs."""
    regions = [common.Region("ref", 0, 5000)] * 3
    model = StubModel(preferred=7)
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "probs.npzstore")
        rem = prediction.run_prediction(out, None, regions, model, FakeEncoder(), 500, 100, batch_size="auto",
                                        bam_workers=2)
        assert rem == []
        n_windows = sum(s[0] for s in model.shapes)
        assert all(s[0] <= 7 and s[1:] == (500, 10) for s in model.shapes)
        assert sum(1 for s in model.shapes if s[0] == 7) >= len(model.shapes) - 1      # only the last batch is short
        with datastore.DataStore(out, "r") as ds:
            # identical regions give identical sample names: the store keeps each name once
            assert ds.n_samples == n_windows // 3
            name = sorted(ds.sample_registry)[0]
            smp = ds.load_sample(name)
            assert smp.label_probs.shape == (500, 5)
            assert smp.labels.shape == (500,) and smp.labels.dtype == np.uint8 and int(smp.labels[0]) == 3
        assert 1 <= model.max_in_flight <= model.depth and model.in_flight == 0


def test_model_archive_roundtrip():
    """An archive written here has the reference's layout: model/weights.pt + model/meta.pkl whose pickles name
    medaka.models.model_from_dict / medaka.features.CountsFeatureEncoder / medaka.labels.HaploidLabelScheme
    (medaka/datastore.py:51-175, medaka/training.py:83-96), and loads back through the restricted unpickler."""
    import pickletools
    import tarfile
    from oracle import synth
    sd = synth.synth_state_dict(1)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "model_pt.tar.gz")
        meta = {"model_function": {"type": "GRUModel", "kwargs": {"num_features": 10}},
                "feature_encoder": {"type": "CountsFeatureEncoder", "kwargs": {"normalise": "total"}},
                "label_scheme": "HaploidLabelScheme"}
        datastore.ModelStoreTGZ.write(path, sd, meta)
        with tarfile.open(path) as tar:
            raw = tar.extractfile("model/meta.pkl").read()
        strings = {arg for op, arg, _ in pickletools.genops(raw) if isinstance(arg, str)}
        assert {"medaka.models", "model_from_dict", "medaka.features", "CountsFeatureEncoder", "medaka.labels",
                "HaploidLabelScheme", "functools", "partial"} <= strings
        assert not any(x.startswith("medaka_b200") for x in strings)
        with datastore.ModelStoreTGZ(path) as ms:
            assert ms.get_meta("label_scheme").symbols == "*ACGT"
            assert ms.get_meta("feature_encoder").to_dict()["kwargs"]["normalise"] == "total"
            assert ms.model_kwargs() == {"type": "GRUModel", "kwargs": {"num_features": 10}}
            w = ms._unpack()._weights
            assert sorted(w) == sorted(sd)
            assert np.array_equal(w["linear.bias"].numpy(), sd["linear.bias"])
            out = os.path.join(d, "probs.npzstore")
            ms.copy_meta(out)
            with datastore.DataStore(out, "r") as ds:
                assert ds.get_meta("feature_encoder").normalise == "total"
                assert type(ds.get_meta("label_scheme")).__name__ == "HaploidLabelScheme"


def test_reference_pickled_meta_loads(golden_dir):
    """meta items pickled by the REAL reference classes (tests/golden/make_meta_golden.py): the v2 form (partial of
    model_from_dict) and the legacy form (partial of build_model_torch); anything else is refused."""
    import pickle
    g = np.load(os.path.join(golden_dir, "ref_meta.npz"))
    m = datastore.ref_loads(g["v2"].tobytes())
    enc = m["feature_encoder"]
    assert (enc.normalise, tuple(enc.dtypes), enc.min_mapq, enc.sym_indels) == ("fwd_rev", ("r9", "r10"), 3, True)
    assert enc.feature_vector_length == 20 and enc.pileup_source is None
    assert sorted(enc.feature_indices) == sorted([("r9", True), ("r9", False), ("r10", True), ("r10", False)])
    assert m["label_scheme"].symbols == "*ACGT" and m["label_scheme"].num_classes == 5
    with tempfile.TemporaryDirectory() as d:
        for key, expect in (("v2", {"num_features": 20, "num_classes": 5, "gru_size": 128}),
                            ("legacy", {"num_features": 10, "num_classes": 5, "gru_size": 128})):
            ms = datastore.ModelStoreTGZ(os.path.join(d, "x.tar.gz"))
            ms._meta = datastore.ref_loads(g[key].tobytes())
            ms._weights = {}
            assert ms.model_kwargs() == {"type": "GRUModel", "kwargs": expect}
    # a read-level (rl_) archive: LatentSpaceLSTM + ReadAlignmentFeatureEncoder as the reference pickles them
    rl = datastore.ref_loads(g["read_level"].tobytes())
    fe = rl["feature_encoder"]
    assert type(fe).__name__ == "ReadAlignmentFeatureEncoder"
    assert (fe.max_reads, fe.include_dwells, fe.min_mapq, fe.row_per_read, fe.normalise) == (80, True, 2, False, None)
    assert fe.feature_vector_length == 5
    ms = datastore.ModelStoreTGZ("unused.tar.gz")
    ms._meta, ms._weights = rl, {}
    kw = ms.model_kwargs()
    assert kw["type"] == "LatentSpaceLSTM" and kw["kwargs"]["use_dwells"] is True and kw["kwargs"]["lstm_size"] == 128
    evil = pickle.dumps(os.system)
    with pytest.raises(pickle.UnpicklingError):
        datastore.ref_loads(evil)


class FakeH5(object):
    """Stands in for the h5py module (this image has no libhdf5): records every dataset with the arguments it was
    created with, keeps files in memory by name."""

    files = {}

    class _DS(object):
        def __init__(self, data, **kw):
            self.data, self.kw = data, kw

        def __getitem__(self, key):
            assert key == ()
            return self.data

    class _Group(object):
        def __init__(self, store, prefix):
            self.store, self.prefix = store, prefix

        def keys(self):
            return sorted({k[len(self.prefix):].split("/")[0] for k in self.store if k.startswith(self.prefix)})

        def __iter__(self):
            return iter(self.keys())

        def __getitem__(self, k):
            return self.store[self.prefix + k]

    class File(object):
        def __init__(self, filename, mode):
            if mode == "w" or filename not in FakeH5.files:
                if mode == "r":
                    raise FileNotFoundError(filename)
                FakeH5.files[filename] = {}
            self.store, self.closed = FakeH5.files[filename], False

        def create_dataset(self, loc, data=None, **kw):
            assert loc not in self.store
            self.store[loc] = FakeH5._DS(np.array(data), **kw)

        def __setitem__(self, loc, value):
            assert loc not in self.store
            self.store[loc] = FakeH5._DS(value)

        def __getitem__(self, loc):
            if loc in self.store:
                return self.store[loc]
            if any(k.startswith(loc + "/") for k in self.store):
                return FakeH5._Group(self.store, loc + "/")
            raise KeyError(loc)

        def __contains__(self, loc):
            return loc in self.store or any(k.startswith(loc + "/") for k in self.store)

        def __delitem__(self, loc):
            del self.store[loc]

        def flush(self):
            pass

        def close(self):
            self.closed = True


def test_hdf_backend_layout():
    """The HDF container writes what `medaka sequence` / `medaka vcf` read (medaka/datastore.py:263-336): datasets at
    samples/data/<name>/<field>, every ndarray field gzip-1 compressed, ref_name as a plain string, the registry and
    the meta items as np.bytes_ pickles under samples/registry and meta/<key>; a sample is written once."""
    import pickle
    FakeH5.files.clear()
    n = 30
    s = common.Sample("ctg", None, np.full(n, 2, dtype=np.uint8), None, _positions(n, 7),
                      np.random.rand(n, 5).astype(np.float32), np.arange(n))
    with datastore.DataStore("out.hdf", "a", h5=FakeH5) as ds:
        ds.set_meta(datastore.as_reference_meta({"label_scheme": "HaploidLabelScheme"})["label_scheme"], "label_scheme")
        ds.write_sample(s)
        ds.write_sample(s)
        assert ds.n_samples == 1
    store = FakeH5.files["out.hdf"]
    base = "samples/data/{}/".format(s.name)
    assert sorted(k[len(base):] for k in store if k.startswith(base)) == ["depth", "label_probs", "labels", "positions",
                                                                         "ref_name"]
    for field in ("depth", "label_probs", "labels", "positions"):
        assert store[base + field].kw == {"compression": "gzip", "compression_opts": 1}, field
    assert store[base + "ref_name"].kw == {} and store[base + "ref_name"].data == "ctg"
    assert isinstance(store["samples/registry"].data, np.bytes_)
    assert pickle.loads(store["samples/registry"].data) == {s.name}
    blob = store["meta/label_scheme"].data
    assert isinstance(blob, np.bytes_) and b"medaka.labels" in bytes(blob) and b"medaka_b200" not in bytes(blob)
    with datastore.DataStore("out.hdf", "r", h5=FakeH5) as ds:
        assert ds.sample_registry == {s.name}
        back = ds.load_sample(s.name)
        assert back.ref_name == "ctg" and np.array_equal(back.label_probs, s.label_probs)
        assert np.array_equal(back.labels, s.labels) and np.array_equal(back.positions, s.positions)
        assert type(ds.get_meta("label_scheme")).__name__ == "HaploidLabelScheme"


WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from medaka_b200 import common, prediction
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
regs = [common.Region("c%d" % i, 0, 100000 + 37000 * ((i * 7) % 11)) for i in range(23)]
long, rem = prediction.triage_regions(regs, 10000, 250000, 1000)
mine = prediction.shard_regions(long, world)[rank]
# weights travel once from rank 0 (the bench does this over NCCL); everything else is rank-local
w = torch.arange(10, dtype=torch.float32) if rank == 0 else torch.zeros(10)
dist.broadcast(w, src=0)
names = [None] * world
dist.all_gather_object(names, [(r.ref_name, r.start, r.end) for r in mine])
sizes = torch.tensor([float(sum(r.size for r in mine))])
dist.all_reduce(sizes, op=dist.ReduceOp.MAX)
if rank == 0:
    flat = sorted(x for n in names for x in n)
    ok = flat == sorted((r.ref_name, r.start, r.end) for r in long) and float(w.sum()) == 45.0
    print("RESULT", json.dumps({{"ok": ok, "n": len(flat), "max_load": float(sizes[0])}}))
dist.destroy_process_group()
"""


def test_region_sharding_world_size_2_gloo():
    """N>1 path on CPU: 2 ranks over gloo shard the regions disjointly and completely."""
    with tempfile.TemporaryDirectory() as d:
        script = os.path.join(d, "worker.py")
        with open(script, "w") as fh:
            fh.write(WORKER.format(root=ROOT))
        r = subprocess.run(
            [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
             "--master-addr", "127.0.0.1", "--master-port", "29533", script],
            capture_output=True, text=True, timeout=240)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        assert line, r.stdout + r.stderr
        import json
        res = json.loads(line[0][7:])
        assert res["ok"] and res["n"] > 23
