#!/bin/bash
# timing experiments on the ping-pong recurrent kernel: each check in its own process under a timeout.  usage: tools/gpu_pp.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
log=gpurun_out/pp_$tag.log
: > $log
run() { echo "== $2 $3" >> $log; timeout $1 python tools/diag.py --check $2 --arg $3 2>&1 | grep -v "^Traceback\|^  File\|^    " | tail -n 12 >> $log; echo "rc=$?" >> $log; }
for f in 0 1 2 3 4 7 8 16 32 63; do run 300 pp_trace 1111,10000,$f; done
python tools/pp_summary.py $log
timeout 900 python -m pytest tests/test_pileup.py tests/test_variants.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_$tag.log
cat gpurun_out/pytest_$tag.log
