"""Exactly 3 warm-up steps then ONE full LM train step (for ncu launch lists: skip the warm-up launches with -s)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slamkit_b200.lm import B200UnitLM, LMConfig, B200AdamW
from slamkit_b200 import _lib
m = B200UnitLM(LMConfig(), device="cuda:0", max_batch=8, max_seq=1024, seed=0)
opt = B200AdamW(m)
g = torch.Generator().manual_seed(1)
ids = torch.randint(2, 502, (8, 1024), generator=g); ids[:, 0] = 1
ids = ids.cuda(); labels = ids.clone()
lib = _lib.load()
n0 = lib.sk_launch_count()
m.forward_backward(ids, labels, num_items_in_batch=8192.0); opt.step()
torch.cuda.synchronize()
per_step = lib.sk_launch_count() - n0
for _ in range(2):
    m.forward_backward(ids, labels, num_items_in_batch=8192.0); opt.step()
torch.cuda.synchronize()
print("launches per step", per_step, flush=True)
m.forward_backward(ids, labels, num_items_in_batch=8192.0); opt.step()
torch.cuda.synchronize()
