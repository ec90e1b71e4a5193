#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
grep -E "^\{" gpurun_out/bench_n$N.json | tail -1 > gpurun_out/bench_n$N.line; python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n$N.line").read())
print("LM", d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
s = d.get("secondary"); print("HUBERT", s and s["value"], s and s["ms_per_batch"])
PY
tail -3 gpurun_out/bench_n$N.err
