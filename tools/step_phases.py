"""Where one LM train step spends its time, from the backward-pass events (sk_lm_set_backward_events): forward + head,
every 2 layers of the backward pass, tail (embedding backward, reduction wait, clip + AdamW).  Run alone (N=1) or under
torchrun (N>1) to see which phase the data-parallel reduction stretches."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from slamkit_b200 import _lib as L
from slamkit_b200.lm import B200UnitLM, LMConfig
from slamkit_b200.trainer import B200Trainer

model = B200UnitLM(LMConfig(), device=str(dev), max_batch=8, max_seq=1024, seed=0)
tr = B200Trainer(model, lr=1e-3, min_lr=5e-5, warmup_steps=100, total_steps=17625, max_grad_norm=0.5)
nl = model.config.n_layers
if world == 1:      # no GradSync events at N=1: install timing events ourselves
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(nl + 1)]
    for e in evs:
        e.record()
    arr = (C.c_void_p * (nl + 1))(*[C.c_void_p(e.cuda_event) for e in evs])
    L.check(model.lib.sk_lm_set_backward_events(model._h, arr, nl + 1))
else:
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(nl + 1)]
    for e in evs:
        e.record()
    arr = (C.c_void_p * (nl + 1))(*[C.c_void_p(e.cuda_event) for e in evs])
    L.check(model.lib.sk_lm_set_backward_events(model._h, arr, nl + 1))
    tr.sync.events = evs            # GradSync waits on the same (timing-enabled) events
g = torch.Generator().manual_seed(rank)
ids = torch.randint(2, 502, (8, 1024), generator=g).to(dev)
mb = [{"input_ids": ids, "labels": ids, "n_items": 8192, "n_tokens": 8192}]
for _ in range(6):
    tr.train_step(mb)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
acc = None
R = 5
for _ in range(R):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    tr.train_step(mb)
    e.record()
    torch.cuda.synchronize()
    t = [s.elapsed_time(evs[l]) for l in range(nl, -1, -1)] + [s.elapsed_time(e)]   # event nl fires first, event 0 last
    acc = t if acc is None else [a + b for a, b in zip(acc, t)]
t = [a / R for a in acc]
out = [f"rank {rank}/{world}: step {t[-1]:.3f} ms | forward+head+final-norm bwd {t[0]:.3f} |"]
prev = t[0]
for k in range(2, nl + 1, 2):
    out.append(f"{t[k] - prev:.3f}")
    prev = t[k]
out.append(f"| tail (embed bwd, reduce wait, clip+AdamW) {t[-1] - t[nl]:.3f} | backward layers total {t[nl] - t[0]:.3f}")
for r in range(world):
    if r == rank:
        print(" ".join(out), flush=True)
    if world > 1:
        dist.barrier()
if world > 1:
    dist.destroy_process_group()
