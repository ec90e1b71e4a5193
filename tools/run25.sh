#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2; grep -E "^FAILED|^ERROR|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | head -10
for m in 2 2; do timeout 600 python tools/hubert_time.py > gpurun_out/hubert_time_$m.json 2> gpurun_out/hubert_time.err; python - <<PY
import json
s = json.load(open("gpurun_out/hubert_time_$m.json"))
print("HUBERT", round(s["value"],3), round(s["ms_per_batch"],2), round(s["roofline"]["achieved"],1), round(s["roofline"]["frac"],4), s["roofline"]["breakdown_ms"])
PY
done
timeout 600 python tools/lm_step_time.py 2>&1 | tail -1
