#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python tools/hbm_bench.py > $O/r02c_hbm_bench.json 2> $O/r02c_hbm_bench.err; tail -n 22 $O/r02c_hbm_bench.json | head -12; tail -n 3 $O/r02c_hbm_bench.err
timeout 1200 ncu --set full --clock-control none -k regex:"plp_" -c 40 -o $O/r02c_prof_plp python tools/hbm_bench.py --n 100000 > $O/r02c_ncu_p.log 2>&1
ncu -i $O/r02c_prof_plp.ncu-rep --page raw --csv > $O/r02c_prof_plp_raw.csv 2>/dev/null; rm -f $O/r02c_prof_plp.ncu-rep
python tools/hbm_summary.py $O/r02c_prof_plp_raw.csv "ncu plp" | tail -n 14
