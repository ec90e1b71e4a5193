#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2; grep -E "^FAILED|^ERROR|^E  " gpurun_out/pytest_gpu.log | head -30
timeout 300 python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1; cat gpurun_out/attn_bench.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("LM", d["value"], d["ms_per_step"], d["roofline"]["breakdown_ms"])
s = d["secondary"]; print("HUBERT", s["value"], s["ms_per_batch"], s["roofline"]["achieved"], s["roofline"]["frac"], s["roofline"]["breakdown_ms"])
PY
tail -3 gpurun_out/bench.err
