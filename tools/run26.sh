#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_lm.py tests/test_gpu_dpo.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2; grep -E "^FAILED|^ERROR|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | head -10
timeout 300 python tools/elementwise_bench.py 2>&1 | tail -8 | tee gpurun_out/elementwise_bench.txt
timeout 600 python tools/lm_step_time.py 2>&1 | tail -1
