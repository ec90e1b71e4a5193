"""Times the full-size LM train step (fwd+bwd+optimiser) and prints a per-kernel breakdown via torch profiler-free
CUDA events around the three phases."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slamkit_b200.lm import B200UnitLM, LMConfig, B200AdamW

m = B200UnitLM(LMConfig(), device="cuda:0", max_batch=8, max_seq=1024, seed=0)
opt = B200AdamW(m)
g = torch.Generator().manual_seed(1)
ids = torch.randint(2, 502, (8, 1024), generator=g); ids[:, 0] = 1
ids = ids.cuda(); labels = ids.clone()
for _ in range(3):
    m.forward_backward(ids, labels, num_items_in_batch=8192.0); opt.step()
torch.cuda.synchronize()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
N = 10
tf = tb = to = 0.0
for _ in range(N):
    evs[0].record(); m.forward(ids, labels=labels, num_items_in_batch=8192.0)
    evs[1].record(); m.forward_backward(ids, labels, num_items_in_batch=8192.0)
    evs[2].record(); opt.step(); evs[3].record()
    torch.cuda.synchronize()
    tf += evs[0].elapsed_time(evs[1]); tb += evs[1].elapsed_time(evs[2]); to += evs[2].elapsed_time(evs[3])
print(f"fwd {tf/N:.2f} ms | fwd+bwd {tb/N:.2f} ms | opt {to/N:.2f} ms | step(fwd+bwd+opt) {(tb+to)/N:.2f} ms"
      f" -> {8192/((tb+to)/N)*1e3:.0f} tok/s ; loss {float(m.stats[0]):.4f}")
