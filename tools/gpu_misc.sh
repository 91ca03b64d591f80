#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r02v_bench.json').read().strip().splitlines()[-1]); print('value %.3e e2e %.3e ratio %.3f'%(d['value'], d['e2e']['value'], d['e2e']['value']/d['value']), d['config']['batches_in_flight'])" || tail -5 gpurun_out/r02v_bench.err
