#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2; grep -E "^FAILED|^ERROR|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | head -10
timeout 150 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("LM", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["breakdown_ms"], d["roofline"]["frac"], d["roofline"]["step_frac_of_peak"], d["clocks"])
s = d["secondary"]; print("HUBERT", s["value"], s["ms_per_batch"], s["e2e"]["value"], s["roofline"]["achieved"], s["roofline"]["frac"], s["roofline"]["breakdown_ms"])
PY
tail -3 gpurun_out/bench.err
