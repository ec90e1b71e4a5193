"""Cost of the backward-pass events (sk_lm_set_backward_events) at N=1: 20 back-to-back train steps with and without them."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slamkit_b200 import _lib as L
from slamkit_b200.lm import B200UnitLM, LMConfig
from slamkit_b200.trainer import B200Trainer
dev = torch.device("cuda", 0)
model = B200UnitLM(LMConfig(), device=str(dev), max_batch=8, max_seq=1024, seed=0)
tr = B200Trainer(model, lr=1e-3, min_lr=5e-5, warmup_steps=100, total_steps=17625, max_grad_norm=0.5)
nl = model.config.n_layers
ids = torch.randint(2, 502, (8, 1024)).to(dev)
mb = [{"input_ids": ids, "labels": ids, "n_items": 8192, "n_tokens": 8192}]
def run(tag):
    for _ in range(5):
        tr.train_step(mb)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        tr.train_step(mb)
    e.record(); torch.cuda.synchronize()
    print(f"{tag}: {s.elapsed_time(e) / 20:.3f} ms/step", flush=True)
run("no events")
for timing in (False, True):
    evs = [torch.cuda.Event(enable_timing=timing) for _ in range(nl + 1)]
    for ev in evs:
        ev.record()
    arr = (C.c_void_p * (nl + 1))(*[C.c_void_p(ev.cuda_event) for ev in evs])
    L.check(model.lib.sk_lm_set_backward_events(model._h, arr, nl + 1))
    run(f"25 events per step (timing={timing})")
L.check(model.lib.sk_lm_set_backward_events(model._h, None, 0))
run("no events again")
