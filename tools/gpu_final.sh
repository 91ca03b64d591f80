#!/bin/bash
# round-end evidence: bench line (with CPU baseline), NT=2 shape, ncu launch list, ncu full capture of the hot kernels
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
timeout 600 python bench.py --windows 2368 --no-cpu-baseline > gpurun_out/bench2368_$tag.json 2>> gpurun_out/bench_$tag.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 16 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_l_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rec_tc|gemm_tc|head_kernel" -s 4 -c 4 -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_f_$tag.log 2>&1
tail -c 400 gpurun_out/bench_$tag.json; tail -c 300 gpurun_out/bench2368_$tag.json
