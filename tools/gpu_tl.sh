#!/bin/bash
# lane schedule experiments: usage tools/gpu_tl.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
log=gpurun_out/tl_$tag.log
: > $log
for spec in pp,2368,4000,8 pp,1184,4000,8 one,1184,4000,8 one,2368,4000,8; do
    echo "== $spec" >> $log
    timeout 300 python tools/diag.py --check timeline --arg $spec 2>&1 | tail -n 2 >> $log
done
cat $log
