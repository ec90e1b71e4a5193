#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hubert.py tests/test_gpu_cli.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2; grep -E "^FAILED|^ERROR|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | head -10
for d in 0 2 0 2; do echo "SK_CONV0_DBG=$d"; SK_CONV0_DBG=$d timeout 600 python tools/hubert_time.py --steps 8 > gpurun_out/hubert_time_d$d.json 2> gpurun_out/hubert_time.err; python - <<PY
import json
s = json.load(open("gpurun_out/hubert_time_d$d.json"))
print("HUBERT", round(s["ms_per_batch"],2), s["roofline"]["breakdown_ms"], round(s["roofline"]["frac"],3))
PY
done
