#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2; grep -E "^FAILED|^ERROR|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | head -10
for v in 0 1 0 1; do echo "SK_PDL=$v"; SK_PDL=$v timeout 600 python tools/lm_step_time.py 2>&1 | tail -1; done
for v in 0 1; do echo "SK_PDL=$v"; SK_PDL=$v timeout 600 python tools/hubert_time.py --steps 12 2>/dev/null | python -c "import json,sys; s=json.loads(sys.stdin.read()); print('HUBERT', round(s['value'],3), round(s['ms_per_batch'],2))"; done
