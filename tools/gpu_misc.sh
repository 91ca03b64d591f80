#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $O/r02f_pytest_gpu.log 2>&1; grep -v "^  File" $O/r02f_pytest_gpu.log | tail -12
timeout 300 python tools/rl_bench.py > $O/r02f_rl_bench.json 2> $O/r02f_rl_bench.err; cat $O/r02f_rl_bench.json
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02f_bench_cfg2.json 2> $O/r02f_bench_cfg2.err; python -c "
import json; d=json.loads(open('$O/r02f_bench_cfg2.json').read().strip().splitlines()[-1]); print('cfg2 value %.3e e2e %.3e ms %.1f'%(d['value'], d['e2e']['value'], d['ms_per_step']), d['clocks'], round(d['roofline']['frac'],3), d['roofline']['full_wave_solo']['frac'])"
timeout 600 python bench.py --config 4 --steps 4 --warmup 3 --no-cpu-baseline > $O/r02f_bench_cfg4.json 2> $O/r02f_bench_cfg4.err; python -c "
import json; d=json.loads(open('$O/r02f_bench_cfg4.json').read().strip().splitlines()[-1]); print('cfg4 value %.3e e2e %.3e'%(d['value'], d['e2e']['value']), d['e2e'].get('variant_decode_ms_per_step'))"
