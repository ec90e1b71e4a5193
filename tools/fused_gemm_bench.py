"""Times the fused-epilogue linears of the LM step against the unfused kernel chains they replace (CUDA events, L2 flushed).
SK_GEMM_EW=4|8 forces the epilogue warp count for an A/B."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slamkit_b200 import ops
from slamkit_b200.lm import rope_tables

dev = "cuda:0"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)


M, d, F, H, KVH = 8192, 896, 4864, 14, 2
x, wg, wu, wd = rnd(M, d), rnd(F, d, scale=0.03), rnd(F, d, scale=0.03), rnd(d, F, scale=0.03)
wgu_b, wgu = ops.block_gate_up(wg, wu), torch.cat([wg, wu], 0).contiguous()
t_f = timeit(lambda: ops.linear_swiglu_fwd(x, wgu_b))
gu = ops.gemm(x, wgu)
t_u = timeit(lambda: ops.gemm(x, wgu)) + timeit(lambda: ops.swiglu_fwd(gu))
print(f"gate/up + SwiGLU fwd     fused {t_f:7.1f} us   unfused chain {t_u:7.1f} us   ({2.0 * M * 2 * F * d / t_f / 1e6:6.1f} TF/s fused)")
dy = rnd(M, d)
gu_b, _ = ops.linear_swiglu_fwd(x, wgu_b)
t_f = timeit(lambda: ops.linear_swiglu_bwd(dy, wd, gu_b))
dact = ops.gemm(dy, wd, b_mn=True)
t_u = timeit(lambda: ops.gemm(dy, wd, b_mn=True)) + timeit(lambda: ops.swiglu_bwd(gu, dact))
print(f"down dgrad + SwiGLU bwd  fused {t_f:7.1f} us   unfused chain {t_u:7.1f} us   ({2.0 * M * F * d / t_f / 1e6:6.1f} TF/s fused)")
N = (H + 2 * KVH) * 64
wq, bq = rnd(N, d, scale=0.03), rnd(N)
cos, sin = rope_tables(10000.0, 64, 2048)
cos, sin = cos.to(dev), sin.to(dev)
t_f = timeit(lambda: ops.linear_rope(x, wq, bq, cos, sin, 1024, (H + KVH) * 64))
q = ops.gemm(x, wq, bias=bq)
t_u = timeit(lambda: ops.gemm(x, wq, bias=bq)) + timeit(lambda: ops.rope_(q, cos, sin, 1024, H + KVH, 64))
print(f"qkv + bias + RoPE        fused {t_f:7.1f} us   unfused chain {t_u:7.1f} us   ({2.0 * M * N * d / t_f / 1e6:6.1f} TF/s fused)")
