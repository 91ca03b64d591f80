#!/bin/bash
# round-end check on one B200: smoke, the GPU test suite, the two bench arms the driver runs.  usage: tools/gpu_final.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 1200 python -m pytest tests -x -q -m gpu > $O/${tag}_pytest_gpu.log 2>&1; tail -n 3 $O/${tag}_pytest_gpu.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/${tag}_bench_reference.json 2> $O/${tag}_bench_reference.err; tail -c 300 $O/${tag}_bench_reference.json
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${tag}_bench_cfg2.json 2> $O/${tag}_bench_cfg2.err; python -c "
import json; d=json.loads(open('$O/${tag}_bench_cfg2.json').read().strip().splitlines()[-1]); print('cfg2 value %.3e e2e %.3e ms %.1f'%(d['value'], d['e2e']['value'], d['ms_per_step']), d['clocks'], round(d['roofline']['frac'],3), round(d['roofline']['full_wave_solo']['frac'],3), d['cpu_baseline']['value'])"
timeout 600 python bench.py --config 4 --steps 4 --warmup 3 --no-cpu-baseline > $O/${tag}_bench_cfg4.json 2> $O/${tag}_bench_cfg4.err; python -c "
import json; d=json.loads(open('$O/${tag}_bench_cfg4.json').read().strip().splitlines()[-1]); print('cfg4 value %.3e e2e %.3e'%(d['value'], d['e2e']['value']), d.get('variant_decode'))" || tail -5 $O/${tag}_bench_cfg4.err
