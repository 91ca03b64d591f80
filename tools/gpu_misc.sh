#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_read_level.py -x -q -m gpu > gpurun_out/misc_pytest.log 2>&1; grep -v "^  File" gpurun_out/misc_pytest.log | tail -25
