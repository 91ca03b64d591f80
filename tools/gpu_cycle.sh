#!/bin/bash
# one GPU visit: parity tests, cycle trace of the recurrent kernel, the bench line.  usage: tools/gpu_cycle.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_$tag.log
timeout 300 python tools/diag.py --check rec_trace --arg 1111,10000 > gpurun_out/trace_$tag.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
cat gpurun_out/pytest_$tag.log; tail -c 600 gpurun_out/bench_$tag.json
