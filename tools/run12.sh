#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hubert.py -m gpu -q -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_hub.log
grep -E "passed|failed" gpurun_out/pytest_hub.log | tail -2; grep -E "^FAILED|^ERROR|^E  " gpurun_out/pytest_hub.log | head -20
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("LM", d["value"], d["ms_per_step"], d["roofline"]["breakdown_ms"])
s = d["secondary"]; print("HUBERT", s["value"], s["ms_per_batch"], s["roofline"]["achieved"], s["roofline"]["frac"], s["roofline"]["breakdown_ms"])
PY
tail -3 gpurun_out/bench.err
