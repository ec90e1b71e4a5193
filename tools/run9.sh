#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("LM", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["step_frac_of_peak"], d["roofline"]["breakdown_ms"], d["clocks"], d["cpu_baseline"])
s = d["secondary"]; print("HUBERT", s["value"], s["ms_per_batch"], s["roofline"]["achieved"], s["roofline"]["breakdown_ms"], s.get("cpu_baseline"))
PY
tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json | cut -c1-300
# launch list of one full step with the final kernel set
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2700 -c 800 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-hubert > gpurun_out/ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt 2>&1; cat gpurun_out/launch_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd_dkdv -s 2 -c 1 -o gpurun_out/prof_attn_tc_dkdv python tools/attn_bench.py > gpurun_out/ncu_a1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd_dq -s 2 -c 1 -o gpurun_out/prof_attn_tc_dq python tools/attn_bench.py > gpurun_out/ncu_a2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd -s 2 -c 1 -o gpurun_out/prof_attn_tc_fwd python tools/attn_bench.py > gpurun_out/ncu_a3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -s 60 -c 6 -o gpurun_out/prof_gemm2 python tools/lm_step_time.py > gpurun_out/ncu_g.log 2>&1
ls -la gpurun_out | tail -12
