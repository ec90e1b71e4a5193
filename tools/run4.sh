#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2; grep -E "^FAILED|^ERROR|^E " gpurun_out/pytest_gpu.log | head -20
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
# hubert launch list + conv0 full capture
cat > /tmp/hub_once.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch
from slamkit_b200.feature_extractor import HubertB200Config, HubertB200FeatureExtractor, random_params
cfg = HubertB200Config(); fe = HubertB200FeatureExtractor(cfg, random_params(cfg, 0), max_batch=16, max_samples=480000)
w = (0.1 * torch.randn(16, 480000)).clamp(-1, 1).cuda()
for _ in range(3): fe.units_device(w, None)
torch.cuda.synchronize()
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 140 --csv --log-file gpurun_out/hubert_launches.csv python /tmp/hub_once.py > gpurun_out/ncu_hub.log 2>&1
python tools/summarize_launches.py gpurun_out/hubert_launches.csv > gpurun_out/hubert_launch_summary.txt 2>&1; cat gpurun_out/hubert_launch_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv0_apply -s 2 -c 1 -o gpurun_out/prof_conv0 python /tmp/hub_once.py > gpurun_out/ncu_conv0.log 2>&1
tail -2 gpurun_out/ncu_conv0.log
