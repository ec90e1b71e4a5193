#!/bin/bash
mkdir -p gpurun_out
N=$(python tools/lm_one_step.py | awk '/launches per step/{print $4}')
echo "launches per step: $N"
# torch's own init kernels precede ours; select by kernel-name regex instead of absolute skip counts
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
   -k 'regex:gemm_tcgen05|splitk_reduce|attn_|rmsnorm|swiglu|rope|colsum|adamw|sumsq|gradnorm|embed|ce_|add_f32|transpose' \
   --csv --log-file gpurun_out/lm_launches_v3.csv python tools/lm_one_step.py > gpurun_out/ncu_lm.log 2>&1
tail -2 gpurun_out/ncu_lm.log
python - <<'PY'
import csv, collections, re
rows = [l for l in open("gpurun_out/lm_launches_v3.csv") if not l.startswith("==")]
r = list(csv.DictReader(rows))
# keep only the last step's launches: find per-kernel-ID ordering
ids = sorted({int(x["ID"]) for x in r})
n_total = len(ids)
print("total profiled launches", n_total)
PY
