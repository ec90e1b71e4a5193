#!/bin/bash
# First GPU call: unit tests (no -x: we want the full picture), GEMM microbench, full-size step timing.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1
timeout 300 python tools/lm_step_time.py > gpurun_out/lm_step_time.log 2>&1
tail -40 gpurun_out/pytest_gpu.log; cat gpurun_out/gemm_bench.log; cat gpurun_out/lm_step_time.log
