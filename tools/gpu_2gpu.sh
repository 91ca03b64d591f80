#!/bin/bash
# two-GPU evidence: bench (weak scaling, weights broadcast over NCCL) and the pipeline bench.  run with gpurun --gpus 2
mkdir -p gpurun_out
P=29517
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
tail -c 700 gpurun_out/r02_bench_2gpu.json; tail -n 3 gpurun_out/r02_bench_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) tools/pipeline_bench.py --mb 40 > gpurun_out/r02_pipeline_2gpu.json 2> gpurun_out/r02_pipeline_2gpu.err
tail -n 2 gpurun_out/r02_pipeline_2gpu.json; tail -n 3 gpurun_out/r02_pipeline_2gpu.err
