#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --config 4 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_cfg4.json 2> gpurun_out/r02e_bench_cfg4.err; python -c "
import json; d=json.loads(open('gpurun_out/r02e_bench_cfg4.json').read().strip().splitlines()[-1]); print('cfg4 value %.3e e2e %.3e'%(d['value'], d['e2e']['value']), {k:v for k,v in d['e2e'].items() if k not in ('value','unit')})" || tail -5 gpurun_out/r02e_bench_cfg4.err
timeout 600 python tools/pipeline_bench.py --mb 100 --profile > gpurun_out/r02e_pipeline_100mb.json 2> gpurun_out/r02e_pipeline_100mb.err; cat gpurun_out/r02e_pipeline_100mb.json; grep "stage seconds" gpurun_out/r02e_pipeline_100mb.err | cut -c1-1200
