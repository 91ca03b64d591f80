#!/bin/bash
# one GPU call: GPU test suite, bench A/B.  usage: tools/gpu_round.sh <tag> [pytest-args]
tag=${1:-x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
bash tools/gpu_bench.sh $tag
