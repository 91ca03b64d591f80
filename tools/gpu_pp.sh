#!/bin/bash
# bring-up / timing of the ping-pong recurrent kernel: each check in its own process under a timeout.  usage: tools/gpu_pp.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
log=gpurun_out/pp_$tag.log
: > $log
run() { echo "== $2 $3" >> $log; timeout $1 python tools/diag.py --check $2 --arg $3 2>&1 | grep -v "^Traceback\|^  File\|^    " | tail -n 12 >> $log; echo "rc=$?" >> $log; }
run 240 pp 37,130,10
run 240 pp 1217,33,20
run 300 rec_timing pp,1111,10000
run 300 pp_trace 1111,10000
cat $log | cut -c1-3500
