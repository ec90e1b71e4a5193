#!/bin/bash
# ncu launch list (time + DRAM bytes per launch) of LM train steps, then the per-layer timeline of the last step
mkdir -p gpurun_out
N=$(python tools/lm_one_step.py | awk '/launches per step/{print $4}')
echo "launches per step: $N"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
   -k 'regex:gemm_tcgen05|splitk_reduce|attn_|rmsnorm|swiglu|rope|colsum|adamw|sumsq|gradnorm|embed|ce_|add_f32|transpose|seg_bounds|lmhead' \
   --csv --log-file gpurun_out/lm_launches.csv python tools/lm_one_step.py > gpurun_out/ncu_lm.log 2>&1
tail -2 gpurun_out/ncu_lm.log
python tools/summarize_launches.py gpurun_out/lm_launches.csv > gpurun_out/lm_launches_summary.txt
python tools/lm_layer_timeline.py gpurun_out/lm_launches.csv $N | tee gpurun_out/lm_layer_timeline.txt
