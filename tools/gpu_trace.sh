#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
timeout 300 python tools/diag.py --check rec_trace --arg 1111,10000 > gpurun_out/trace_$tag.log 2>&1
tail -c 300 gpurun_out/trace_$tag.log
