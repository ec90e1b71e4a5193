"""cli/extract_features.py -- drop-in for the reference entry point (cli/extract_features.py:71-99): audio directory ->
features jsonl, one `{"units": [...], "duration": [...], "file_name": ...}` line per file, files processed in
descending-length batches (unit ids depend on batch composition).

    python cli/extract_features.py data_path=<dir> ext=wav out_path=<features.jsonl> batch_size=16 \
        tokeniser=unit_hubert_25 tokeniser.feature_extractor_type=hubert_b200

Under torchrun every rank takes whole batches round-robin, writes `<out_path>.rank{r}`, and rank 0 merges them back into
`<out_path>` in global batch order.  Files are decoded by `num_workers` threads ahead of the GPU and each
batch's host-to-device copy runs on a copy stream from pinned memory (BatchPrefetcher).
`+synthetic_weights=true` builds a seeded random mHuBERT-geometry extractor (no checkpoint reachable offline)."""
import json
import logging
import os
import sys
from glob import iglob

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from slamkit_b200.audio_io import audio_num_frames, load_audio  # noqa: E402
from slamkit_b200.config import load_config, require  # noqa: E402

logger = logging.getLogger(__name__)
LAST_STATS = {}     # throughput of the most recent main() call (extraction loop only: no model construction)


def build_tokeniser(cfg, device: str, max_batch=None):
    from slamkit_b200.feature_extractor import HubertB200Config, HubertB200FeatureExtractor, random_params
    from slamkit_b200.integration import hubert_b200_from_cfg
    from slamkit_b200.tokeniser import B200UnitTokeniser
    t = cfg.tokeniser
    if t.tokeniser_type != "unit":
        raise ValueError(f"Unknown tokeniser type: {t.tokeniser_type}")          # audio_tokeniser.py:121
    if t.feature_extractor_type not in ("hubert", "hubert_b200"):
        raise ValueError(f"Unknown feature extractor type: {t.feature_extractor_type}")   # audio_tokeniser.py:104
    fe_args = dict(t.feature_extractor)
    max_batch = max_batch or cfg.batch_size
    fe = None
    if t.params.get("load_fe", True):
        if cfg.get("synthetic_weights", False):
            hc = HubertB200Config(layer=fe_args["layer"], n_units=fe_args["num_units"])
            fe = HubertB200FeatureExtractor(hc, random_params(hc, seed=0), device=device, max_batch=max_batch,
                                            max_samples=16000 * 30, load_config_only=fe_args.get("load_config_only", False))
        else:
            fe = hubert_b200_from_cfg(**fe_args, device=device, max_batch=max_batch)
    p = t.params
    return B200UnitTokeniser(fe, dedup=p.dedup, bos_eos_token_id=p.get("bos_eos_token_id", 1), pad_token_id=p.pad_token_id,
                             num_units=p.get("num_units") or fe_args["num_units"], load_fe=p.get("load_fe", True))


class BatchPrefetcher:
    """What `DataLoader(num_workers=4, collate_fn=pad_wav_collate)` + `.to(device)` do in the reference
    (cli/extract_features.py:86-93), arranged for a GPU that finishes a 64 x 30 s batch in ~65 ms: `num_workers` threads
    decode / resample the files of upcoming batches (the FLAC decoder and numpy release the GIL), a collector thread pads
    each batch into one of `depth` PINNED host buffers and issues its host-to-device copy on a separate CUDA stream, and
    the consumer receives device tensors whose copy overlapped the previous batch's kernels.  Order is the batch order."""

    def __init__(self, batches, sample_rate: int, device: str, num_workers: int = 4, depth: int = 2):
        import queue
        import threading
        from concurrent.futures import ThreadPoolExecutor
        self.batches, self.sr, self.dev = batches, sample_rate, torch.device(device)
        self.pool = ThreadPoolExecutor(max_workers=max(1, num_workers))
        self.q = queue.Queue(maxsize=depth)
        self.free = queue.Queue()
        for _ in range(depth + 1):
            self.free.put(None)                      # pinned buffers are created lazily at the size of the largest batch seen
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        try:
            # decode futures run ahead of the collector by two batches
            futs = [[self.pool.submit(load_audio, f, self.sr) for f, _ in b] for b in self.batches[:2]]
            for i, batch in enumerate(self.batches):
                if i + 2 < len(self.batches):
                    futs.append([self.pool.submit(load_audio, f, self.sr) for f, _ in self.batches[i + 2]])
                wavs = [ft.result() for ft in futs[i]]
                futs[i] = None
                lens = torch.tensor([len(w) for w in wavs])
                B, S = len(wavs), int(lens.max())
                buf = self.free.get()
                if buf is None or buf.numel() < B * S:
                    buf = torch.empty(B * S, dtype=torch.float32).pin_memory()
                host = buf[:B * S].view(B, S)
                host.zero_()
                for r, w in enumerate(wavs):
                    host[r, :len(w)] = w
                with torch.cuda.stream(self.copy_stream):
                    dev_wav = host.to(self.dev, non_blocking=True)
                    dev_lens = lens.to(self.dev, non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(self.copy_stream)
                self.q.put((batch, dev_wav, dev_lens, ready, buf))
            self.q.put(None)
        except BaseException as e:                   # surface loader errors in the consumer
            self.q.put(e)

    def __iter__(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            if isinstance(item, BaseException):
                raise item
            batch, dev_wav, dev_lens, ready, buf = item
            torch.cuda.current_stream().wait_event(ready)
            dev_wav.record_stream(torch.cuda.current_stream())
            dev_lens.record_stream(torch.cuda.current_stream())
            yield batch, dev_wav, dev_lens
            # the pinned buffer is reusable once its copy has executed; the event was recorded right after it
            ready.synchronize()
            self.free.put(buf)


def merge_rank_files(out_path: str, world: int) -> None:
    """Deterministic merge of the per-rank outputs: rank r wrote whole batches r, r + world, ... in order,
    one json line per file; interleave them back into global batch order (= the single-process file order)."""
    recs = [[json.loads(l) for l in open(f"{out_path}.rank{r}")] for r in range(world)]
    n_batches = [json.load(open(f"{out_path}.rank{r}.batches")) for r in range(world)]
    with open(out_path, "a+") as out:
        pos = [0] * world
        for bi in range(max(len(n) for n in n_batches)):
            for r in range(world):
                if bi < len(n_batches[r]):
                    k = n_batches[r][bi]
                    out.writelines(json.dumps(x) + "\n" for x in recs[r][pos[r]:pos[r] + k])
                    pos[r] += k
    for r in range(world):
        os.remove(f"{out_path}.rank{r}")
        os.remove(f"{out_path}.rank{r}.batches")


def main(argv=None):
    cfg = load_config("extract_features", argv if argv is not None else sys.argv[1:])
    require(cfg, "data_path", "out_path")
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    device = f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"
    torch.cuda.set_device(device)
    files = [(f, audio_num_frames(f)) for f in iglob(os.path.join(cfg.data_path, f"**/*.{cfg.ext}"), recursive=True)]
    files.sort(key=lambda x: x[1], reverse=True)          # WavDataset: sort by duration, longest first
    if cfg.data_skip is not None:
        files = files[cfg.data_skip:]
    if cfg.data_take is not None:
        files = files[:cfg.data_take]
    tokeniser = build_tokeniser(cfg, device)
    out_path = cfg.out_path if world == 1 else f"{cfg.out_path}.rank{rank}"
    if os.path.exists(out_path):
        if world > 1:
            os.remove(out_path)                           # rank files are scratch; the merged file keeps append semantics
        else:
            logging.warning(f"{out_path} already exists. Appending to it.")
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    batches = [files[i:i + cfg.batch_size] for i in range(0, len(files), cfg.batch_size)]
    mine = [b for bi, b in enumerate(batches) if bi % world == rank]      # whole batches round-robin (ids depend on batch composition)
    sizes = []
    import time
    t_loop, n_samples = time.perf_counter(), 0
    with open(out_path, "a+") as out_file:
        for batch, wav, lens in BatchPrefetcher(mine, cfg.sample_rate, device, num_workers=cfg.get("num_workers", 4) or 1):
            n_samples += sum(n for _, n in batch)
            reps = tokeniser.audio_represent(wav, lens)
            lines = []
            for (f, _), rep in zip(batch, reps):
                rec = {"units": list(rep["units"]), "duration": list(rep["duration"]), "file_name": f}
                lines.append(json.dumps(rec) + "\n")
            out_file.writelines(lines)
            sizes.append(len(batch))
    dt = time.perf_counter() - t_loop
    LAST_STATS.update({"files": sum(sizes), "seconds": dt, "audio_hours": n_samples / cfg.sample_rate / 3600.0,
                       "audio_hours_per_s": n_samples / cfg.sample_rate / 3600.0 / max(dt, 1e-9)})
    logger.info(f"rank {rank}: {LAST_STATS['files']} files, {LAST_STATS['audio_hours']:.2f} audio-hours in {dt:.2f} s "
                f"({LAST_STATS['audio_hours_per_s']:.2f} audio-hours/s: decode + copy + extraction + jsonl)")
    if world > 1:
        import torch.distributed as dist
        json.dump(sizes, open(out_path + ".batches", "w"))
        if not dist.is_initialized():
            dist.init_process_group("gloo")
        dist.barrier()
        if rank == 0:
            merge_rank_files(cfg.out_path, world)
        dist.barrier()
        dist.destroy_process_group()
        return cfg.out_path
    return out_path


if __name__ == "__main__":
    main()
