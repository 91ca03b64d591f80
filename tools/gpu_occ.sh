#!/bin/bash
# recurrent-kernel step time against the number of busy SMs: usage tools/gpu_occ.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
log=gpurun_out/occ_$tag.log
: > $log
for spec in one,280,4000 one,560,4000 one,1111,4000 one,1184,4000 pp,560,4000 pp,1111,4000 pp,1184,4000 pp,2222,4000 pp,2368,4000; do
  echo "== $spec" >> $log
  timeout 300 python tools/diag.py --check rec_timing --arg $spec,2 2>&1 | tail -n 1 | cut -c1-330 >> $log
done
cat $log
