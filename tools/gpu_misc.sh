#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pipeline.py tests/test_read_level.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/pipeline_bench.py --mb 100 > gpurun_out/r02_pipeline_1gpu_100mb.json 2> gpurun_out/r02_pipeline_100mb.err; cat gpurun_out/r02_pipeline_1gpu_100mb.json
