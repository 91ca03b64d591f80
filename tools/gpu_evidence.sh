#!/bin/bash
# round-2 evidence: bench lines (config 2 with CPU baseline, reference arm, configs 3-5), precision table, HBM kernel
# bench, ncu launch list + full captures.  usage: tools/gpu_evidence.sh <tag> [part ...]   parts: tests bench cfg prec hbm ncu pipe
tag=${1:-x}; shift
parts=${@:-bench cfg prec hbm ncu}
mkdir -p gpurun_out
O=gpurun_out
for part in $parts; do
case $part in
tests)
  timeout 1800 python -m pytest tests -x -q -m gpu > $O/${tag}_pytest_gpu.log 2>&1; tail -n 4 $O/${tag}_pytest_gpu.log ;;
bench)
  timeout 900 python bench.py --steps 20 --warmup 5 > $O/${tag}_bench_cfg2.json 2> $O/${tag}_bench_cfg2.err
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/${tag}_bench_reference.json 2> $O/${tag}_bench_reference.err
  timeout 900 python bench.py --impl reference --steps 1 --warmup 0 --cpu-cols 10000 > $O/${tag}_bench_reference_fullT.json 2> $O/${tag}_bench_reference_fullT.err
  tail -c 600 $O/${tag}_bench_cfg2.json; tail -c 400 $O/${tag}_bench_reference.json ;;
cfg)
  for c in 3 4 5; do
    timeout 900 python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline > $O/${tag}_bench_cfg$c.json 2> $O/${tag}_bench_cfg$c.err
    tail -c 300 $O/${tag}_bench_cfg$c.json
  done ;;
prec)
  timeout 1200 python tools/precision_table.py > $O/${tag}_precision.md 2> $O/${tag}_precision.err; tail -n 5 $O/${tag}_precision.md ;;
hbm)
  timeout 900 python tools/hbm_bench.py --cpu-pileup > $O/${tag}_hbm_bench.json 2> $O/${tag}_hbm_bench.err; tail -n 30 $O/${tag}_hbm_bench.json; tail -n 5 $O/${tag}_hbm_bench.err ;;
ncu)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 40 --csv --log-file $O/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${tag}_ncu_l.log 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"rec_pp|gemm_tc|head_plog" -s 6 -c 4 -o $O/${tag}_prof_fwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${tag}_ncu_f.log 2>&1
  timeout 1200 ncu --set full --clock-control none -k regex:"normalise|decode_kernel|head_plog|stitch_|plp_|vd_" -c 60 -o $O/${tag}_prof_hbm python tools/hbm_bench.py --n 4000000 > $O/${tag}_ncu_h.log 2>&1
  # the reports stay on the box (gpurun_out is capped at 64 MiB): export the raw metric pages, keep the small one
  ncu -i $O/${tag}_prof_fwd.ncu-rep --page raw --csv > $O/${tag}_prof_fwd_raw.csv 2>/dev/null
  ncu -i $O/${tag}_prof_hbm.ncu-rep --page raw --csv > $O/${tag}_prof_hbm_raw.csv 2>/dev/null
  ls -la $O/${tag}_prof_*; rm -f $O/${tag}_prof_hbm.ncu-rep
  [ $(stat -c %s $O/${tag}_prof_fwd.ncu-rep) -gt 30000000 ] && rm -f $O/${tag}_prof_fwd.ncu-rep ;;
pipe)
  timeout 900 python tools/pipeline_bench.py --mb 20 > $O/${tag}_pipeline_1gpu.json 2> $O/${tag}_pipeline_1gpu.err; tail -n 2 $O/${tag}_pipeline_1gpu.json; tail -n 3 $O/${tag}_pipeline_1gpu.err ;;
esac
done
