"""Element-wise / reduction kernels of the LM step timed alone at the bench shape (M = 8192 rows), L2 flushed between
iterations; prints microseconds and the fraction of the measured HBM peak their algorithmic bytes reach."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slamkit_b200 import ops

dev = "cuda:0"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    PEAK = float(PEAK.get("hbm_gbs") or PEAK.get("hbm", {}).get("gbs") or 6590.6)
except Exception:
    PEAK = 6590.6


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


def report(name, us, nbytes):
    print(f"{name:16s} {us:7.1f} us   {nbytes / us / 1e3:7.0f} GB/s   {nbytes / us / 1e3 / PEAK:5.2f} of HBM peak", flush=True)


# fixed cost of one launch + event pair in this harness (same kernel on 8 rows)
_x = torch.randn(8, 896, device=dev).to(torch.bfloat16); _w = torch.ones(896, device=dev).to(torch.bfloat16)
print(f"(harness floor: {timeit(lambda: ops.rmsnorm_fwd(_x, _w, 1e-6)):.1f} us per single-launch op)")
M, D, F, QKV = 8192, 896, 4864, 1152
bf = torch.bfloat16
x = torch.randn(M, D, device=dev).to(bf); dy = torch.randn(M, D, device=dev).to(bf); dres = torch.randn(M, D, device=dev).to(bf)
w = torch.ones(D, device=dev).to(bf); dw = torch.zeros(D, device=dev).to(bf)
y, rstd = ops.rmsnorm_fwd(x, w, 1e-6)
report("rmsnorm_fwd", timeit(lambda: ops.rmsnorm_fwd(x, w, 1e-6)), 2 * M * D * 2)
report("rmsnorm_bwd", timeit(lambda: ops.rmsnorm_bwd(dy, x, w, rstd, dres, dw, True)), 4 * M * D * 2)
gu = torch.randn(M, 2 * F, device=dev).to(bf); dact = torch.randn(M, F, device=dev).to(bf)
report("swiglu_fwd", timeit(lambda: ops.swiglu_fwd(gu)), 3 * M * F * 2)
report("swiglu_bwd", timeit(lambda: ops.swiglu_bwd(gu, dact)), 5 * M * F * 2)
qkv = torch.randn(M, QKV, device=dev).to(bf)
from slamkit_b200.lm import rope_tables
cos, sin = rope_tables(1e6, 64, 1024); cos, sin = cos.to(dev), sin.to(dev)
report("rope", timeit(lambda: ops.rope_(qkv, cos, sin, 1024, 16, 64)), 2 * M * 16 * 64 * 2)
out = torch.zeros(QKV, device=dev).to(bf)
report("colsum(qkv)", timeit(lambda: ops.colsum(qkv, out, True)), M * QKV * 2)
