#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_lm.py -m gpu -q -x -k "attn or lm" -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2; grep -E "^FAILED|^ERROR|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | head -10
timeout 300 python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1; cat gpurun_out/attn_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_tc --csv --log-file gpurun_out/attn_launches.csv python tools/attn_bench.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/attn_launches.csv | head -8
