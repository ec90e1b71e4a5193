#!/bin/bash
mkdir -p gpurun_out
./tools/micro/ffma2_bench > gpurun_out/ffma2_bench.txt 2>&1; cat gpurun_out/ffma2_bench.txt
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2; grep -E "^FAILED|^ERROR|^E  " gpurun_out/pytest_gpu.log | head -30
for sk in 0 1 0 1; do echo "SK_STREAMK=$sk"; SK_STREAMK=$sk timeout 600 python tools/lm_step_time.py 2>&1 | tail -1; done
