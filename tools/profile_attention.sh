#!/bin/bash
# ncu --set full captures (with source-level stall attribution) of the three tcgen05 attention kernels at the LM shape
mkdir -p gpurun_out
for k in attn_tc_fwd_kernel attn_tc_bwd_dkdv attn_tc_bwd_dq; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_$k python tools/attn_bench.py > gpurun_out/ncu_$k.log 2>&1
  tail -2 gpurun_out/ncu_$k.log
done
python tools/attn_bench.py
