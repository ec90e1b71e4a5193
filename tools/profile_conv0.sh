#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/hub_once.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch
from slamkit_b200.feature_extractor import HubertB200Config, HubertB200FeatureExtractor, random_params
cfg = HubertB200Config(); fe = HubertB200FeatureExtractor(cfg, random_params(cfg, 0), max_batch=16, max_samples=480000)
w = (0.1 * torch.randn(16, 480000)).clamp(-1, 1).cuda()
for _ in range(2): fe.units_device(w, None)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv0_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_conv0_tc python /tmp/hub_once.py > gpurun_out/ncu_conv0.log 2>&1
tail -1 gpurun_out/ncu_conv0.log
