#!/bin/bash
# bench A/B: usage tools/gpu_bench.sh <tag> [extra bench args]
tag=${1:-x}; shift
mkdir -p gpurun_out
for mode in one pp; do
  timeout 900 python bench.py --no-cpu-baseline --rec-mode $mode "$@" > gpurun_out/bench_${tag}_$mode.json 2> gpurun_out/bench_${tag}_$mode.err
  tail -c 2500 gpurun_out/bench_${tag}_$mode.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$mode', 'value %.3e e2e %.3e ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d['roofline']['solo_stage_ms'], d['roofline']['timed_region_stage_ms'], d['clocks'])" || tail -5 gpurun_out/bench_${tag}_$mode.err
done
