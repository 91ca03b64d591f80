#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/rl_bench.py > gpurun_out/r02_rl_bench.json 2> gpurun_out/r02_rl_bench.err; cat gpurun_out/r02_rl_bench.json; tail -3 gpurun_out/r02_rl_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rl_|head_kernel" -c 40 --csv --log-file gpurun_out/r02_rl_launches.csv python tools/rl_bench.py --cpu-windows 1 > /dev/null 2>&1; grep -v "^==" gpurun_out/r02_rl_launches.csv | awk -F'","' 'NR>1{print $5, $(NF)}' | tail -12
