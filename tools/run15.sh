#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm" -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_gemm.log
grep -E "passed|failed" gpurun_out/pytest_gemm.log | tail -2; grep -E "^FAILED|^ERROR|^E  " gpurun_out/pytest_gemm.log | head -30
timeout 600 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log
