#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_lm.py -m gpu -q -x -k "attention or packed or seg_bounds or lm_" -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2; grep -E "^FAILED|^ERROR|^E  " gpurun_out/pytest_gpu.log | cut -c1-400 | head -12
timeout 300 python tools/attn_bench.py 2>&1 | tail -6
