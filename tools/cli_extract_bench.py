"""End-to-end throughput of cli/extract_features.py on a synthetic WAV directory:
writes N clips of 30 s @ 16 kHz (16-bit PCM) to a temp dir, runs the CLI's main() (threaded decode, pinned prefetch, GPU
extraction, jsonl writing) and reports audio-hours per second of the whole call (model construction excluded by timing a
first small call separately).

    python tools/cli_extract_bench.py [n_clips=256] [batch_size=64] [num_workers=8]"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cli import extract_features
from slamkit_b200.audio_io import write_wav

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 256
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 64
workers = int(sys.argv[3]) if len(sys.argv) > 3 else 8
d = tempfile.mkdtemp()
g = torch.Generator().manual_seed(0)
base = (0.1 * torch.randn(480000, generator=g)).clamp(-1, 1)
for i in range(n_clips):
    write_wav(os.path.join(d, f"clip{i:05d}.wav"), torch.roll(base, i * 977))
args = [f"data_path={d}", "ext=wav", f"batch_size={bs}", f"num_workers={workers}", "tokeniser.feature_extractor_type=hubert_b200",
        "+synthetic_weights=true"]
t0 = time.perf_counter()
extract_features.main(args + [f"out_path={d}/warm.jsonl", f"data_take={bs}"])        # builds the model, warms the kernels
t_warm = time.perf_counter() - t0
t0 = time.perf_counter()
out = extract_features.main(args + [f"out_path={d}/f.jsonl"])
dt = time.perf_counter() - t0
n = sum(1 for _ in open(out))
hours = n * 30.0 / 3600.0
print(json.dumps({"metric": "cli/extract_features.py audio-hours/s (whole main(): file scan, threaded WAV decode, pinned prefetch, "
                            "extraction, dedup, jsonl)", "clips": n, "batch_size": bs, "num_workers": workers,
                  "seconds": dt, "audio_hours_per_s": hours / dt, "first_call_seconds_incl_model_build": t_warm,
                  "extraction_loop": dict(extract_features.LAST_STATS)}))
