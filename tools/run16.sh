#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2; grep -E "^FAILED|^ERROR|^E  " gpurun_out/pytest_gpu.log | head -30
timeout 600 python tools/hubert_time.py > gpurun_out/hubert_time.json 2> gpurun_out/hubert_time.err; python - <<'PY'
import json
s = json.load(open("gpurun_out/hubert_time.json"))
print("HUBERT", s["value"], s["ms_per_batch"], s["roofline"]["achieved"], s["roofline"]["frac"], s["roofline"]["breakdown_ms"])
PY
tail -3 gpurun_out/hubert_time.err
timeout 600 python tools/lm_step_time.py 2>&1 | tail -8
