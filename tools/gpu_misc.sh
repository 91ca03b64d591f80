#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_read_level.py -x -q -m gpu > $O/misc_pytest.log 2>&1; grep -v "^  File" $O/misc_pytest.log | tail -25
timeout 200 python tools/rl_bench.py > $O/r02g_rl_bench.json 2> $O/r02g_rl_bench.err; cat $O/r02g_rl_bench.json; tail -3 $O/r02g_rl_bench.err
