#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd_dkdv -s 2 -c 1 -f -o gpurun_out/prof_dkdv_v2 python tools/attn_bench.py > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd_dq -s 2 -c 1 -f -o gpurun_out/prof_dq_v2 python tools/attn_bench.py > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log
