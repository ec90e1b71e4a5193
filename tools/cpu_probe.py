import os, time, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try: print("cgroup cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e: print("no cgroup v2 cpu.max", e)
os.system("lscpu | grep -E 'Model name|Socket|Thread|Core' | head -5")
a = torch.randn(1024, 896).bfloat16(); w = torch.randn(9728, 896).bfloat16()
for th in (8, 16, 32, 64, 128):
    if th > os.cpu_count(): break
    torch.set_num_threads(th)
    torch.nn.functional.linear(a, w)
    t = time.time()
    for _ in range(5): torch.nn.functional.linear(a, w)
    print("threads", th, "bf16 linear ms", (time.time() - t) / 5 * 1e3, flush=True)
