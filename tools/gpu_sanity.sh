#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
cat > /tmp/small_fwd.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from medaka_b200 import models
from oracle import synth
for (B, T, F) in ((37, 130, 10), (20, 40, 20)):
    sd = synth.synth_state_dict(1, num_features=F)
    m = models.GRUModel(num_features=F); m.load_state_dict(sd)
    out = m.forward_arrays(synth.synth_features(B, T, F, seed=2), want_logits=True, want_labels=True)
    print(B, T, F, float(out.probs.sum()), out.labels[:1, :8])
    m.close()
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/small_fwd.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck.log
tail -5 gpurun_out/smoke.log; tail -15 gpurun_out/memcheck.log
