#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rl_|head_kernel" -s 8 -c 8 --csv --log-file $O/r02_rl_launches.csv python tools/rl_bench.py --windows 64 --reads 30 --cpu-windows 0 --tc-only > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"rl_conv17_tc|rl_lstm_tc" -s 3 -c 2 -o $O/r02_prof_rl python tools/rl_bench.py --windows 64 --reads 30 --cpu-windows 0 --tc-only > $O/r02_ncu_rl.log 2>&1
ncu -i $O/r02_prof_rl.ncu-rep --page raw --csv > $O/r02_prof_rl_raw.csv 2>/dev/null; rm -f $O/r02_prof_rl.ncu-rep; ls -la $O/r02_prof_rl_raw.csv
