"""Output seam: storing `Sample`s (positions, label_probs, depth[, features]) and model archives.

Keeps the reference's logical layout (medaka/datastore.py:178-360): samples live under
``samples/data/<sample.name>/<field>``, the set of written names under ``samples/registry``,
pickled meta items under ``meta/<key>``; a sample already in the registry is not written twice
(coarse resume, datastore.py:277-299); writes go through one background thread.

Two interchangeable containers implement that layout:
  * HDF5 through h5py, byte-compatible with what ``medaka sequence`` / ``medaka vcf`` read
    (used whenever h5py is importable);
  * ``NpzDirStore`` - one ``.npz`` per sample in a directory - for hosts without libhdf5
    (this build image has neither h5py nor libhdf5).
``DataStore(filename, mode)`` picks by availability and file suffix.
"""
import contextlib
import functools
import io
import json
import os
import struct
import pickle
import sys
import tarfile
import threading
import types
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from medaka_b200 import common

try:  # pragma: no cover - not present in the build image
    import h5py
except ImportError:  # noqa
    h5py = None


def _to_numpy(x):
    if hasattr(x, "detach"):
        return x.detach().cpu().numpy()
    return x


# ------------------------------------------------------------------------------------------------------------------
# Pickles that cross the fence to the reference.  medaka stores OBJECTS in its archives and output files:
# ``meta.pkl`` of a model archive and the ``meta/*`` datasets of a consensus HDF hold ``model_function`` = a
# functools.partial of a medaka.models function, ``feature_encoder`` = a medaka.features.CountsFeatureEncoder and
# ``label_scheme`` = a medaka.labels.HaploidLabelScheme (medaka/datastore.py:135-157,171-175; medaka/training.py:83-96;
# read back at medaka/prediction.py:117-131 and, from the HDF, at medaka/stitch.py / variant.py via
# ``index.metadata['label_scheme']``).  Reading: a restricted Unpickler maps exactly those globals to this package's
# classes and refuses everything else (a model archive is downloaded data).  Writing: the same objects are pickled
# under the reference's module paths, so that an unmodified `medaka sequence` / `medaka vcf` can load what we wrote.
def _ref_model_from_dict(d, time_steps=None, device=0):
    from medaka_b200 import models
    return models.model_from_dict(d, time_steps=time_steps, device=device)


def _ref_build_model_torch(feature_len, num_classes, gru_size=128, classify_activation='softmax', time_steps=None,
                           device=0):
    from medaka_b200 import models
    return models.build_model_torch(feature_len, num_classes, gru_size=gru_size, time_steps=time_steps, device=device)


def _ref_globals():
    from medaka_b200 import features, labels
    import collections
    return {
        ('functools', 'partial'): functools.partial,
        ('collections', 'OrderedDict'): collections.OrderedDict,
        ('collections', 'defaultdict'): collections.defaultdict,
        ('builtins', 'list'): list, ('builtins', 'tuple'): tuple, ('builtins', 'dict'): dict, ('builtins', 'set'): set,
        ('builtins', 'object'): object,
        ('copyreg', '_reconstructor'): __import__('copyreg')._reconstructor,
        ('medaka.models', 'model_from_dict'): _ref_model_from_dict,
        ('medaka.models', 'build_model_torch'): _ref_build_model_torch,
        ('medaka.features', 'CountsFeatureEncoder'): features.CountsFeatureEncoder,
        ('medaka.features', 'ReadAlignmentFeatureEncoder'): features.ReadAlignmentFeatureEncoder,
        ('medaka.labels', 'HaploidLabelScheme'): labels.HaploidLabelScheme,
        # archives written by this package
        ('medaka_b200.features', 'CountsFeatureEncoder'): features.CountsFeatureEncoder,
        ('medaka_b200.features', 'ReadAlignmentFeatureEncoder'): features.ReadAlignmentFeatureEncoder,
        ('medaka_b200.labels', 'HaploidLabelScheme'): labels.HaploidLabelScheme,
        ('medaka_b200.datastore', '_ref_model_from_dict'): _ref_model_from_dict,
        ('medaka_b200.datastore', '_ref_build_model_torch'): _ref_build_model_torch,
    }


class _RefUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        try:
            return _ref_globals()[module, name]
        except KeyError:
            raise pickle.UnpicklingError(
                "refusing to unpickle {}.{}: medaka_b200 loads GRU / LatentSpaceLSTM models with a Counts- or "
                "ReadAlignmentFeatureEncoder and a HaploidLabelScheme only".format(module, name))


def ref_loads(data):
    """Unpickle a meta item written by the reference (or by this package)."""
    return _RefUnpickler(io.BytesIO(bytes(data))).load()


@contextlib.contextmanager
def _reference_module_names():
    """While active, this package's stand-ins answer to the reference's module paths, so that pickle records
    ``medaka.labels HaploidLabelScheme`` etc. (pickle checks that sys.modules[module].name is the object)."""
    from medaka_b200 import features, labels
    fakes = {
        'medaka': types.ModuleType('medaka'),
        'medaka.models': types.ModuleType('medaka.models'),
        'medaka.features': types.ModuleType('medaka.features'),
        'medaka.labels': types.ModuleType('medaka.labels'),
    }
    renamed = [(_ref_model_from_dict, 'medaka.models', 'model_from_dict'),
               (_ref_build_model_torch, 'medaka.models', 'build_model_torch'),
               (features.CountsFeatureEncoder, 'medaka.features', 'CountsFeatureEncoder'),
               (features.ReadAlignmentFeatureEncoder, 'medaka.features', 'ReadAlignmentFeatureEncoder'),
               (labels.HaploidLabelScheme, 'medaka.labels', 'HaploidLabelScheme')]
    saved_mods = {k: sys.modules.get(k) for k in fakes}
    saved_names = [(o, o.__module__, o.__qualname__, o.__name__) for o, _, _ in renamed]
    try:
        for k, m in fakes.items():
            sys.modules[k] = m
        for obj, mod, name in renamed:
            obj.__module__, obj.__qualname__, obj.__name__ = mod, name, name
            setattr(fakes[mod], name, obj)
        yield
    finally:
        for obj, mod, qn, nm in saved_names:
            obj.__module__, obj.__qualname__, obj.__name__ = mod, qn, nm
        for k, m in saved_mods.items():
            if m is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = m


_ref_dump_lock = threading.Lock()


def ref_dumps(obj):
    """Pickle a meta item the way the reference would have written it (protocol 2-compatible globals)."""
    with _ref_dump_lock, _reference_module_names():
        return pickle.dumps(obj, protocol=4)


def as_reference_meta(meta):
    """Normalise this package's plain-dict meta ({'type','kwargs'} dicts, a label-scheme name) into the objects the
    reference stores: a partial of model_from_dict, an encoder object, a label-scheme object."""
    from medaka_b200 import features, labels
    out = dict(meta)
    mf = out.get('model_function')
    if isinstance(mf, dict):
        out['model_function'] = functools.partial(_ref_model_from_dict, mf)
    fe = out.get('feature_encoder')
    if isinstance(fe, dict):
        cls = features.ReadAlignmentFeatureEncoder if fe.get('type') == 'ReadAlignmentFeatureEncoder' else \
            features.CountsFeatureEncoder
        out['feature_encoder'] = cls(**fe.get('kwargs', {}))
    ls = out.get('label_scheme')
    if isinstance(ls, str):
        if ls != 'HaploidLabelScheme':
            raise NotImplementedError("label scheme {} is not implemented".format(ls))
        out['label_scheme'] = labels.HaploidLabelScheme()
    return out


class _NpzBackend(object):
    """Directory container: <root>/samples/<quoted name>.mds, <root>/meta/<key>.pkl, registry.pkl.

    A sample file is a JSON header (field -> dtype, shape, byte offset; strings inline) followed by the raw array bytes:
    written with one ``write`` per field, no compression and no checksumming, so the writer thread keeps up with the
    engine (the reference's HDF5 gzip-1 datasets, medaka/datastore.py:323-329, and numpy's zip container both cost more
    host time per window than the GPU forward does).  Stores written as .npz by earlier versions still load."""

    _MAGIC = b"MDKS1\n"

    def __init__(self, root, mode):
        self.root = root
        if mode in ("w", "a"):
            os.makedirs(os.path.join(root, "samples"), exist_ok=True)
            os.makedirs(os.path.join(root, "meta"), exist_ok=True)
        elif not os.path.isdir(root):
            raise FileNotFoundError(root)

    @staticmethod
    def _fname(name, ext=".mds"):
        return name.replace("/", "%2F").replace(":", "%3A") + ext

    def write_fields(self, name, fields):
        header, blobs, off = {}, [], 0
        for k, v in fields.items():
            if isinstance(v, str):
                header[k] = {"str": v}
                continue
            a = np.ascontiguousarray(v)
            if a.dtype.kind == "U":
                a = np.char.encode(a)
            if a.dtype.names is not None:
                descr = [[n, a.dtype[n].str] for n in a.dtype.names]
            else:
                descr = a.dtype.str
            header[k] = {"dtype": descr, "shape": list(a.shape), "offset": off, "nbytes": int(a.nbytes)}
            blobs.append(a)
            off += (a.nbytes + 63) // 64 * 64
        head = json.dumps(header).encode()
        tmp = os.path.join(self.root, "samples", self._fname(name) + ".tmp")
        with open(tmp, "wb", buffering=0) as fh:
            fh.write(self._MAGIC + struct.pack("<I", len(head)) + head)
            pos = 0
            for a in blobs:
                fh.write(memoryview(a).cast("B") if a.nbytes else b"")
                pad = (a.nbytes + 63) // 64 * 64 - a.nbytes
                if pad:
                    fh.write(b"\0" * pad)
                pos += a.nbytes + pad
        os.replace(tmp, os.path.join(self.root, "samples", self._fname(name)))

    def read_fields(self, name):
        path = os.path.join(self.root, "samples", self._fname(name))
        if not os.path.exists(path):          # a store written by an earlier version
            with np.load(os.path.join(self.root, "samples", self._fname(name, ".npz")), allow_pickle=False) as z:
                return {k: (z[k].item() if z[k].ndim == 0 and z[k].dtype.kind in "US" else z[k]) for k in z.files}
        with open(path, "rb") as fh:
            raw = fh.read()
        if raw[:len(self._MAGIC)] != self._MAGIC:
            raise IOError("{} is not a sample file".format(path))
        n = struct.unpack_from("<I", raw, len(self._MAGIC))[0]
        base = len(self._MAGIC) + 4
        header = json.loads(raw[base:base + n].decode())
        base += n
        out = {}
        for k, h in header.items():
            if "str" in h:
                out[k] = h["str"]
                continue
            dt = np.dtype([tuple(x) for x in h["dtype"]]) if isinstance(h["dtype"], list) else np.dtype(h["dtype"])
            count = int(np.prod(h["shape"])) if h["shape"] else 1
            a = np.frombuffer(raw, dtype=dt, count=count, offset=base + h["offset"]).reshape(h["shape"]).copy()
            out[k] = a
        return out

    def sample_names(self):
        d = os.path.join(self.root, "samples")
        return {f[:-4].replace("%2F", "/").replace("%3A", ":") for f in os.listdir(d) if f.endswith((".mds", ".npz"))}

    def write_blob(self, path, obj):
        with open(os.path.join(self.root, path.replace("/", os.sep) + ".pkl"), "wb") as fh:
            fh.write(ref_dumps(obj))

    def read_blob(self, path):
        with open(os.path.join(self.root, path.replace("/", os.sep) + ".pkl"), "rb") as fh:
            return ref_loads(fh.read())

    def close(self):
        pass


class _H5Backend(object):
    """HDF5 container with the reference's exact dataset paths and compression choices (medaka/datastore.py:263-329).

    ``h5`` is the h5py module (tests pass a stand-in that records the calls: this image has no libhdf5)."""

    def __init__(self, filename, mode, h5=None):
        self.h5 = h5 if h5 is not None else h5py
        self.fh = self.h5.File(filename, mode)

    def write_fields(self, name, fields):
        for k, v in fields.items():
            loc = "samples/data/{}/{}".format(name, k)
            if isinstance(v, np.ndarray):
                # numpy arrays of unicode go in as bytes (datastore.py:283-286)
                if v.dtype.kind == 'U':
                    v = np.char.encode(v)
                # _write_dataset (datastore.py:323-329): every ndarray field is gzip-1 compressed
                self.fh.create_dataset(loc, data=v, compression="gzip", compression_opts=1)
            else:
                self.fh[loc] = v        # ref_name: a plain string dataset

    def read_fields(self, name):
        g = self.fh["samples/data/{}".format(name)]
        out = {}
        for k in g:
            v = g[k][()]
            out[k] = v.decode() if isinstance(v, bytes) else v
        return out

    def sample_names(self):
        return set(self.fh["samples/data"].keys()) if "samples/data" in self.fh else set()

    def write_blob(self, path, obj):
        # _write_pickled (datastore.py:331-336): np.string_(pickle.dumps(obj)), objects under the reference's module
        # paths so that `medaka sequence` can unpickle index.metadata['label_scheme'] etc.
        if path in self.fh:
            del self.fh[path]
        self.fh[path] = np.bytes_(ref_dumps(obj))
        self.fh.flush()

    def read_blob(self, path):
        return ref_loads(self.fh[path][()])

    def close(self):
        self.fh.close()


class DataStore(object):
    """Read and write `Sample`s (reference interface: medaka/datastore.py:178-360)."""

    _meta_group_ = 'meta'
    _sample_path_ = 'samples/data'
    _sample_registry_path_ = 'samples/registry'

    def __init__(self, filename, mode='r', h5=None):
        self.filename = filename
        self.mode = mode
        self.logger = common.get_named_logger('DataStre')
        h5 = h5 if h5 is not None else h5py
        use_h5 = h5 is not None and not str(filename).endswith(".npzstore")
        self._backend = _H5Backend(filename, mode, h5) if use_h5 else _NpzBackend(filename, mode)
        self.write_executor = ThreadPoolExecutor(1)
        self.write_futures = []
        self._sample_registry = None
        self._lock = threading.Lock()

    def __enter__(self):
        return self

    def __exit__(self, *args):
        if self.mode != 'r':
            self.write_executor.shutdown(wait=True)
            for f in self.write_futures:
                f.result()               # surface writer-thread exceptions
            self._backend.write_blob(self._sample_registry_path_, self.sample_registry)
        self.close()

    def close(self):
        self._backend.close()

    def get_meta(self, key):
        try:
            return self._backend.read_blob('{}/{}'.format(self._meta_group_, key))
        except Exception as e:
            self.logger.debug("Could not load {} from {}. {}.".format(key, self.filename, e))

    def set_meta(self, obj, key):
        self._backend.write_blob('{}/{}'.format(self._meta_group_, key), obj)

    @property
    def sample_registry(self):
        if self._sample_registry is None:
            try:
                self._sample_registry = set(self._backend.read_blob(self._sample_registry_path_))
            except Exception:
                self._sample_registry = set(self._backend.sample_names())
        return self._sample_registry

    @property
    def n_samples(self):
        return len(self.sample_registry)

    def write_sample(self, sample, copy=True):
        """Queue a sample for writing unless its name is already registered.  ``copy=False`` hands the arrays to the
        writer thread as they are: for callers whose arrays are not reused before the store is closed."""
        fields = {f: _to_numpy(getattr(sample, f)) for f in sample._fields if getattr(sample, f) is not None}
        if not any(isinstance(v, np.ndarray) for v in fields.values()):
            self.logger.debug('Not writing sample as it has no data.')
            return
        name = sample.name
        with self._lock:
            if name in self.sample_registry:
                self.logger.debug('Not writing {} as present already'.format(name))
                return
            self._sample_registry.add(name)
        # copy views of pinned / reused buffers before handing them to the writer thread
        if copy:
            fields = {k: (np.array(v) if isinstance(v, np.ndarray) else v) for k, v in fields.items()}
        self.write_futures.append(self.write_executor.submit(self._backend.write_fields, name, fields))

    def load_sample(self, key):
        got = self._backend.read_fields(key)
        s = {x: got.get(x) for x in common.Sample._fields}
        return common.Sample(**s)


class ModelStoreTGZ(object):
    """Model archive: ``model/weights.pt`` + pickled ``model/meta.pkl`` in a tar.gz (medaka/datastore.py:51-175).

    Reads the reference's shipped ``*_model_pt.tar.gz`` archives: ``meta.pkl`` there holds ``model_function`` as a
    functools.partial of medaka.models.model_from_dict / build_model_torch plus pickled medaka FeatureEncoder and
    LabelScheme objects; they are loaded through ``ref_loads`` (restricted to exactly those classes) onto this
    package's stand-ins.  ``write`` produces the same layout (``ref_dumps``), so the archive also loads in the reference.
    A ``model_function`` given as a plain ``{'type', 'kwargs'}`` dict is accepted as well.
    """

    top_level_dir = 'model'

    def __init__(self, filepath):
        self.filepath = filepath
        self._meta = None
        self._weights = None

    def __enter__(self):
        return self

    def __exit__(self, *args):
        pass

    @classmethod
    def write(cls, filepath, state_dict, meta):
        import torch
        with tarfile.open(filepath, "w:gz") as tar:
            buf = io.BytesIO()
            torch.save({k: torch.as_tensor(np.asarray(v)) for k, v in state_dict.items()}, buf)
            for name, data in (("weights.pt", buf.getvalue()), ("meta.pkl", ref_dumps(as_reference_meta(meta)))):
                info = tarfile.TarInfo("{}/{}".format(cls.top_level_dir, name))
                info.size = len(data)
                tar.addfile(info, io.BytesIO(data))

    def _unpack(self):
        if self._meta is None:
            import torch
            with tarfile.open(self.filepath) as tar:
                members = {m.name: m for m in tar.getmembers() if m.isfile()}
                for needed in ("model/weights.pt", "model/meta.pkl"):
                    if needed not in members:
                        raise KeyError("{} is not a model archive: {} missing".format(self.filepath, needed))
                self._meta = ref_loads(tar.extractfile(members["model/meta.pkl"]).read())
                raw = tar.extractfile(members["model/weights.pt"]).read()
                self._weights = torch.load(io.BytesIO(raw), map_location="cpu", weights_only=True)
        return self

    @property
    def meta(self):
        return self._unpack()._meta

    def get_meta(self, key):
        return self.meta[key]

    def copy_meta(self, hdf):
        """Copy metadata to the output store (datastore.py:165-175): objects, under the reference's module paths."""
        with DataStore(hdf, 'a') as ds:
            for k, v in as_reference_meta(self.meta).items():
                ds.set_meta(v, k)

    def model_kwargs(self):
        """The architecture arguments inside ``model_function`` ({'type', 'kwargs'}), whatever its form."""
        mf = self.meta["model_function"]
        if isinstance(mf, dict):
            return mf
        if isinstance(mf, functools.partial):
            if mf.func is _ref_model_from_dict:
                return mf.args[0] if mf.args else mf.keywords['dict']
            if mf.func is _ref_build_model_torch:
                names = ('feature_len', 'num_classes', 'gru_size')
                kw = dict(zip(names, mf.args))
                kw.update({k: v for k, v in mf.keywords.items() if k in names})
                return {'type': 'GRUModel', 'kwargs': {'num_features': kw['feature_len'], 'num_classes': kw['num_classes'],
                                                       'gru_size': kw.get('gru_size', 128)}}
        raise TypeError("unsupported model_function in {}: {!r}".format(self.filepath, mf))

    def load_model(self, device=0, time_steps=None):
        """Build the engine-backed model and load its weights (cf. datastore.py:135-157)."""
        from medaka_b200 import models
        self._unpack()
        model = models.model_from_dict(self.model_kwargs(), time_steps=time_steps, device=device)
        model.load_state_dict(self._weights)
        return model.eval()
