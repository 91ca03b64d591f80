#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02w_bench_default.json 2> gpurun_out/r02w_bench_default.err; python -c "
import json; d=json.loads(open('gpurun_out/r02w_bench_default.json').read().strip().splitlines()[-1]); print('value %.3e e2e %.3e'%(d['value'], d['e2e']['value']), d['dtype'], d['cpu_baseline'], d['steps'], d['warmup'])" || tail -5 gpurun_out/r02w_bench_default.err
