#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
log=gpurun_out/pp_$tag.log
: > $log
run() { echo "== $2 $3" >> $log; timeout $1 python tools/diag.py --check $2 --arg $3 2>&1 | grep -v "^Traceback\|^  File\|^    " | tail -n 12 >> $log; echo "rc=$?" >> $log; }
run 300 rec_timing pp,1111,10000
for f in 0 63; do run 300 pp_trace 1111,10000,$f; done
grep -A1 "== rec_timing" $log | cut -c1-400
python tools/pp_summary.py $log 2>/dev/null | grep -v FAILED
