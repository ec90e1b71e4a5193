"""Timeline of the peer-memory gradient all-reduce inside real LM train steps (2+ GPUs, torchrun): per bucket, when its
gradients became final (READY sent), when the reduce kernel got its first CTA onto an SM, when the peers were ready, and
when the last CTA finished -- relative to the end of the backward pass (the tail range's READY).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/p2p_trace.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
from slamkit_b200 import _lib as L
from slamkit_b200.lm import B200UnitLM, LMConfig
from slamkit_b200.trainer import B200Trainer

model = B200UnitLM(LMConfig(), device=str(dev), max_batch=8, max_seq=1024, seed=0)
tr = B200Trainer(model, lr=1e-3, min_lr=5e-5, warmup_steps=100, total_steps=17625, max_grad_norm=0.5)
assert tr.sync.backend == "p2p"
g = torch.Generator().manual_seed(rank)
ids = torch.randint(2, 502, (8, 1024), generator=g).to(dev)
mb = [{"input_ids": ids, "labels": ids, "n_items": 8192, "n_tokens": 8192}]
trace = torch.zeros(257 * 4, dtype=torch.int64, device=dev)
for _ in range(5):
    tr.train_step(mb)
torch.cuda.synchronize(); dist.barrier()
L.check(model.lib.sk_p2p_set_trace(C.c_void_p(trace.data_ptr())))
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
tr.train_step(mb)
e.record()
torch.cuda.synchronize()
L.check(model.lib.sk_p2p_set_trace(C.c_void_p(0)))
t = trace.cpu().view(257, 4)
nb = len(tr.sync.buckets)
ref = int(t[nb, 0])          # tail READY = backward pass complete
lines = [f"rank {rank}: step {s.elapsed_time(e):.3f} ms; buckets of {tr.sync.buckets[0][2] - tr.sync.buckets[0][1]} elements; times in us relative to the end of backward"]
for k in range(nb + 1):
    r = [(int(x) - ref) / 1e3 for x in t[k]]
    lines.append(f"  slot {k:2d}  ready {r[0]:10.1f}  first CTA {r[1]:10.1f} (+{r[1]-r[0]:7.1f})  peers ready {r[2]:10.1f} (+{r[2]-r[1]:7.1f})  done {r[3]:10.1f} (+{r[3]-r[2]:7.1f})")
w = [(int(x) - ref) / 1e3 for x in t[256][:2]]
lines.append(f"  wait kernel {w[0]:10.1f} -> {w[1]:10.1f}")
for r in range(world):
    if r == rank:
        print("\n".join(lines), flush=True)
    dist.barrier()
dist.destroy_process_group()
