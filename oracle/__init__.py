"""CPU oracle for the medaka consensus-inference hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy / torch-CPU fp32) of the reference's
algorithm for the path SURVEY.md section 8 scopes.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it, and only as the checker or the timed CPU baseline.  Nothing
under ``medaka_b200/`` imports it; the product path fails loudly when the CUDA
library is missing instead of falling back here.

Pinning status (DESIGN.md section "Oracle"):
  * gru_oracle       - pinned to the REAL reference classes: tests/golden/make_golden.py
                       imports /root/reference/medaka/architectures/gru.py (GRUModel) and
                       medaka/models.py (TorchModel.predict_on_batch) in the build container
                       and stores their outputs; tests/test_oracle.py replays them.
                       (The reference's own tests hold no numeric known-answer for the
                       forward pass - SURVEY.md 8c - so these generated vectors are the pin.)
  * features_oracle  - pinned to the reference's own golden vectors
                       (medaka/test/test_counts.py:92-115, :152-174, :298-311, :486-513)
                       and to outputs of the real CountsFeatureEncoder._post_process_pileup.
  * labels_oracle    - pinned to medaka/test/test_labels.py:239-266 and to outputs of the
                       real HaploidLabelScheme.decode_consensus.
  * common_oracle    - pinned to outputs of the real Sample.chunks / Region.split /
                       sliding_window (medaka/common.py:429-453, :712-737, :803-823).
  * pileup_oracle    - restates src/medaka_counts.c:251-361 + src/medaka_bamiter.c:17-45 over
                       plain alignment records; pinned to medaka/test/test_counts.py:298-334
                       (mock_data.py:22-100 reads).  htslib itself (1.14, build.py:10) is not
                       in the tree, so bam_mplp_auto semantics are restated from its documented
                       behaviour and anchored on those goldens only.
"""
