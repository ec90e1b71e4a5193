"""cli/extract_features.py -- drop-in for the reference entry point (cli/extract_features.py:71-99): audio directory ->
features jsonl, one `{"units": [...], "duration": [...], "file_name": ...}` line per file, files processed in
descending-length batches (unit ids depend on batch composition).

    python cli/extract_features.py data_path=<dir> ext=wav out_path=<features.jsonl> batch_size=16 \
        tokeniser=unit_hubert_25 tokeniser.feature_extractor_type=hubert_b200

Under torchrun every rank takes whole batches round-robin and appends to `<out_path>.rank{r}`.
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


def main(argv=None):
    cfg = load_config("extract_features", argv if argv is not None else sys.argv[1:])
    require(cfg, "data_path", "out_path")
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    device = f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"
    files = [(f, audio_num_frames(f)) for f in iglob(os.path.join(cfg.data_path, f"**/*.{cfg.ext}"), recursive=True)]
    files.sort(key=lambda x: x[1], reverse=True)          # WavDataset: sort by duration, longest first
    if cfg.data_skip is not None:
        files = files[cfg.data_skip:]
    if cfg.data_take is not None:
        files = files[:cfg.data_take]
    tokeniser = build_tokeniser(cfg, device)
    out_path = cfg.out_path if world == 1 else f"{cfg.out_path}.rank{rank}"
    if os.path.exists(out_path):
        logging.warning(f"{out_path} already exists. Appending to it.")
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    batches = [files[i:i + cfg.batch_size] for i in range(0, len(files), cfg.batch_size)]
    with open(out_path, "a+") as out_file:
        for bi, batch in enumerate(batches):
            if bi % world != rank:
                continue
            wavs = [load_audio(f, cfg.sample_rate) for f, _ in batch]
            lens = torch.tensor([len(w) for w in wavs])
            wav = torch.nn.utils.rnn.pad_sequence(wavs, batch_first=True, padding_value=0)
            reps = tokeniser.audio_represent(wav, lens)
            lines = []
            for (f, _), rep in zip(batch, reps):
                rec = {"units": list(rep["units"]), "duration": list(rep["duration"]), "file_name": f}
                lines.append(json.dumps(rec) + "\n")
            out_file.writelines(lines)
    return out_path


if __name__ == "__main__":
    main()
