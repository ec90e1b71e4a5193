"""Runs one fused linear a few times (target for ncu): python tools/one_fused.py swiglu_bwd|swiglu_fwd|rope"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slamkit_b200 import ops
from slamkit_b200.lm import rope_tables
dev = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "swiglu_bwd"
def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)
M, d, F, H, KVH = 8192, 896, 4864, 14, 2
x, wg, wu, wd = rnd(M, d), rnd(F, d, scale=0.03), rnd(F, d, scale=0.03), rnd(d, F, scale=0.03)
wgu_b = ops.block_gate_up(wg, wu)
gu_b, _ = ops.linear_swiglu_fwd(x, wgu_b)
dy = rnd(M, d)
cos, sin = rope_tables(10000.0, 64, 2048)
cos, sin = cos.to(dev), sin.to(dev)
N = (H + 2 * KVH) * 64
wq, bq = rnd(N, d, scale=0.03), rnd(N)
for _ in range(5):
    if which == "swiglu_bwd": ops.linear_swiglu_bwd(dy, wd, gu_b)
    elif which == "swiglu_fwd": ops.linear_swiglu_fwd(x, wgu_b)
    else: ops.linear_rope(x, wq, bq, cos, sin, 1024, (H + KVH) * 64)
torch.cuda.synchronize()
