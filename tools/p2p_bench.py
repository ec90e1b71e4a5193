"""Peer-memory all-reduce (csrc/p2p_comm.cu) vs NCCL on the flat bf16 gradient buffer of the LM (358 M elements), timed
alone with CUDA events.  Run under torchrun with >= 2 GPUs:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/p2p_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
from slamkit_b200.p2p import PeerAllReduce

N = 358_293_504 // 8 * 8
g = torch.Generator(device=dev).manual_seed(rank)
buf = (torch.randn(N, device=dev, generator=g) * 0.01).to(torch.bfloat16)
src = buf.clone()
pa = PeerAllReduce(buf)
gathered = [torch.empty_like(src) for _ in range(world)]
dist.all_gather(gathered, src)
want = gathered[0].float()
for x in gathered[1:]:
    want += x.float()
want = want.to(torch.bfloat16)
del gathered


def p2p_once(ctas):
    pa.begin()
    pa.all_reduce(0, N, ctas)
    pa.finish()


def timed(fn, iters=5):
    ts = []
    for _ in range(iters):
        buf.copy_(src)
        torch.cuda.synchronize(); dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


for ctas in [int(c) for c in os.environ.get('P2P_BENCH_CTAS', '16,48,148,296,592,1184').split(',')]:
    t = timed(lambda: p2p_once(ctas))
    ok = bool(torch.equal(buf, want))
    pa.check()
    if rank == 0:
        busbw = 2 * (world - 1) / world * N * 2 / (t * 1e-3) / 1e9
        print(f"p2p  ctas {ctas:5d}: {t:8.3f} ms  bus bandwidth {busbw:7.1f} GB/s  exact {ok}", flush=True)
t = timed(lambda: dist.all_reduce(buf))
if rank == 0:
    print(f"nccl            : {t:8.3f} ms  bus bandwidth {2 * (world - 1) / world * N * 2 / (t * 1e-3) / 1e9:7.1f} GB/s", flush=True)
dist.barrier()
dist.destroy_process_group()
