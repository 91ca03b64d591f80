#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r02u_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/r02u_pytest_gpu.log
