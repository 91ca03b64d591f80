#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_read_matrix.py tests/test_variants.py tests/test_stitch.py tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/misc_pytest.log 2>&1; grep -v "^  File" gpurun_out/misc_pytest.log | tail -25
timeout 600 python tools/diag.py --check e2e_timeline --arg 2778,10000,200,3 2>&1 | tail -1 | cut -c1-300
timeout 600 python tools/diag.py --check e2e_timeline --arg 1111,10000,200,6 2>&1 | tail -1 | cut -c1-300
