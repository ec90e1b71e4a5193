#!/bin/bash
# ncu launch list (time + DRAM bytes per launch) of LM train steps, then the per-layer timeline of the last step
mkdir -p gpurun_out
N=$(python tools/lm_one_step.py | awk '/launches per step/{print $4}')
echo "launches per step: $N"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
   -k 'regex:gemm_tcgen05|splitk_reduce|attn_|rmsnorm|swiglu|rope|colsum|adamw|sumsq|gradnorm|embed|ce_|add_f|transpose|seg_bounds|lmhead' \
   --csv --log-file gpurun_out/lm_launches.csv python tools/lm_one_step.py > gpurun_out/ncu_lm.log 2>&1
tail -2 gpurun_out/ncu_lm.log
python tools/summarize_launches.py gpurun_out/lm_launches.csv > gpurun_out/lm_launches_summary.txt
python tools/lm_layer_timeline.py gpurun_out/lm_launches.csv $N | tee gpurun_out/lm_layer_timeline.txt
# mean DRAM bytes per tcgen05 GEMM launch of the last step -> profiles/r02_lm_gemm_traffic.json (read by bench.py)
python - "$N" <<'PY'
import csv, json, sys
per_step = int(sys.argv[1])
rows = [l for l in open("gpurun_out/lm_launches.csv") if not l.startswith("==")]
by = {}
for r in csv.DictReader(rows):
    d = by.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "b": 0.0})
    if r["Metric Name"].startswith("dram__bytes"):
        d["b"] += float(r["Metric Value"].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r["Metric Unit"]]
allk = [by[i] for i in sorted(by)]
seq = allk[max(i for i, k in enumerate(allk) if "embed_fwd" in k["name"]):]
g = [x["b"] for x in seq if "gemm_tcgen05" in x["name"]]
json.dump({"gemm_dram_bytes_per_launch": sum(g) / len(g), "gemm_launches_per_step": len(g), "gemm_dram_bytes_per_step": sum(g),
           "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the tcgen05 GEMM launches of one [8,1024] step, tools/profile_lm_step.sh"},
          open("gpurun_out/r02_lm_gemm_traffic.json", "w"), indent=1)
print(open("gpurun_out/r02_lm_gemm_traffic.json").read())
PY
