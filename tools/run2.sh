#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_ref.json
# launch list of one profiled step (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 1300 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt 2>&1; cat gpurun_out/launch_summary.txt
# full capture of the widest GEMM (gate|up forward) and the attention kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -s 40 -c 3 -o gpurun_out/prof_gemm \
   python tools/lm_step_time.py > gpurun_out/ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 12 -c 4 -o gpurun_out/prof_attn \
   python tools/lm_step_time.py > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out
