#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_read_level.py -x -q -m gpu > $O/misc_pytest.log 2>&1; grep -v "^  File" $O/misc_pytest.log | tail -8
timeout 200 python tools/rl_bench.py > $O/r02k_rl_bench.json 2> $O/r02k_rl_bench.err; cat $O/r02k_rl_bench.json; tail -2 $O/r02k_rl_bench.err
timeout 200 python tools/rl_bench.py --windows 256 --positions 1000 --reads 30 --cpu-windows 0 --tc-only > $O/r02k_rl_bench_256.json 2>> $O/r02k_rl_bench.err; cat $O/r02k_rl_bench_256.json
