"""Model / output-store metadata as the REFERENCE pickles it (build container only; needs /root/reference).

1. Pickles, with the real reference classes, the three meta items every medaka model archive and consensus HDF carries
   (medaka/training.py:83-96, medaka/datastore.py:165-175): ``model_function`` = functools.partial(
   medaka.models.model_from_dict, cfg), ``feature_encoder`` = medaka.features.CountsFeatureEncoder(...),
   ``label_scheme`` = medaka.labels.HaploidLabelScheme(); plus the legacy form partial(medaka.models.build_model_torch,
   ...).  -> tests/golden/ref_meta.npz (bytes), read back by tests/test_host.py through medaka_b200.datastore.ref_loads.
2. Checks the other direction here, where the reference is importable: what medaka_b200.datastore.ref_dumps writes
   unpickles, with the stock pickle module, into the reference's own classes with the right attributes.
"""
import functools
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import make_golden  # noqa: E402


def main():
    make_golden.install_stubs()
    import numpy as np
    import medaka.features as ref_features
    import medaka.labels as ref_labels
    import medaka.models as ref_models

    cfg = {"type": "GRUModel", "kwargs": {"num_features": 20, "num_classes": 5, "gru_size": 128}}
    meta = {
        "model_function": functools.partial(ref_models.model_from_dict, cfg),
        "feature_encoder": ref_features.CountsFeatureEncoder(normalise='fwd_rev', dtypes=('r9', 'r10'), min_mapq=3,
                                                             sym_indels=True),
        "label_scheme": ref_labels.HaploidLabelScheme(),
    }
    legacy = {
        "model_function": functools.partial(ref_models.build_model_torch, 10, 5, gru_size=128),
        "feature_encoder": ref_features.CountsFeatureEncoder(),
        "label_scheme": ref_labels.HaploidLabelScheme(),
    }
    rl_cfg = {"type": "LatentSpaceLSTM", "kwargs": {"num_classes": 5, "lstm_size": 128, "cnn_size": 128,
                                                      "kernel_sizes": [1, 17], "pooler_type": "mean", "pooler_args": {},
                                                      "use_dwells": True, "bases_alphabet_size": 6,
                                                      "bases_embedding_size": 6, "bidirectional": True}}
    read_level = {
        "model_function": functools.partial(ref_models.model_from_dict, rl_cfg),
        "feature_encoder": ref_features.ReadAlignmentFeatureEncoder(max_reads=80, include_dwells=True, min_mapq=2),
        "label_scheme": ref_labels.HaploidLabelScheme(),
    }
    out = {"v2": np.frombuffer(pickle.dumps(meta, protocol=4), dtype=np.uint8),
           "legacy": np.frombuffer(pickle.dumps(legacy, protocol=2), dtype=np.uint8),
           "read_level": np.frombuffer(pickle.dumps(read_level, protocol=4), dtype=np.uint8)}
    np.savez_compressed(os.path.join(HERE, "ref_meta.npz"), meta="medaka v%s" % __import__('medaka').__version__, **out)

    # ---- the other direction: our pickles in the reference's hands
    from medaka_b200 import datastore
    ours = datastore.ref_dumps(datastore.as_reference_meta({
        "model_function": {"type": "GRUModel", "kwargs": {"num_features": 10}},
        "feature_encoder": {"type": "CountsFeatureEncoder", "kwargs": {"normalise": "total", "dtypes": ("",)}},
        "label_scheme": "HaploidLabelScheme"}))
    got = pickle.loads(ours)
    assert isinstance(got["feature_encoder"], ref_features.CountsFeatureEncoder)
    assert got["feature_encoder"].normalise == "total" and got["feature_encoder"].min_mapq == 1
    assert got["feature_encoder"].feature_indices == ref_features.pileup_counts_norm_indices(("",))
    assert got["feature_encoder"].feature_vector_length == 10 and hasattr(got["feature_encoder"], "logger")
    assert isinstance(got["label_scheme"], ref_labels.HaploidLabelScheme)
    assert got["label_scheme"].num_classes == 5 and got["label_scheme"]._decoding[1] == ('A',)
    mf = got["model_function"]
    assert mf.func is ref_models.model_from_dict and mf.args[0]["type"] == "GRUModel"
    model = mf(time_steps=None)
    assert type(model).__name__ == "GRUModel" and model.gru.hidden_size == 128
    ours_rl = datastore.ref_dumps(datastore.as_reference_meta({
        "model_function": rl_cfg,
        "feature_encoder": {"type": "ReadAlignmentFeatureEncoder", "kwargs": {"max_reads": 60, "include_dwells": False}},
        "label_scheme": "HaploidLabelScheme"}))
    got_rl = pickle.loads(ours_rl)
    fe = got_rl["feature_encoder"]
    assert isinstance(fe, ref_features.ReadAlignmentFeatureEncoder) and fe.max_reads == 60 and fe.include_dwells is False
    assert fe.normalise is None and fe.feature_vector_length == 4 and fe.row_per_read is False
    rl_model = got_rl["model_function"](time_steps=None)
    assert type(rl_model).__name__ == "LatentSpaceLSTM" and rl_model.use_dwells is True and rl_model.lstm_size == 128
    print("reference -> ours: ref_meta.npz written;  ours -> reference: unpickled into", type(got["feature_encoder"]),
          type(got["label_scheme"]), type(model))


if __name__ == "__main__":
    main()
