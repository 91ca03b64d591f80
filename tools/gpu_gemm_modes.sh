#!/bin/bash
mkdir -p gpurun_out
for m in 0 1 2 3; do
  MDK_GEMM_MODE=$m timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.avg,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:gemm_tc -s 1 -c 1 --csv --log-file gpurun_out/gemm_mode$m.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/gemm_mode$m.log 2>&1
  MDK_GEMM_MODE=$m timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('mode $m', d['value'], d['roofline']['stage_ms'], d['clocks']['sm_mhz'])"
  grep -v "^==" gpurun_out/gemm_mode$m.csv | tail -5 | cut -d, -f13-15
done
