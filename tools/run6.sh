#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 3 4 7; do echo "== SK_ATTN_DBG=$d"; SK_ATTN_DBG=$d timeout 120 python tools/attn_bench.py 2>&1 | grep "tcgen05"; done > gpurun_out/attn_dbg.log 2>&1
cat gpurun_out/attn_dbg.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd -s 3 -c 1 -o gpurun_out/prof_attn_tc python tools/attn_bench.py > gpurun_out/ncu_attn_tc.log 2>&1
tail -2 gpurun_out/ncu_attn_tc.log
