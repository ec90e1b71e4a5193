"""cli/preference_alignment_feature_extractor.py -- drop-in for the reference entry point
(cli/preference_alignment_feature_extractor.py:57-82): units for the three clips of every preference triplet.

    python cli/preference_alignment_feature_extractor.py data_path=<triplets.jsonl> out_path=<out.jsonl> [batch_size=8]

Input rows  {"prompt_path": ..., "chosen_path": ..., "rejected_path": ...} (+ any other fields, kept);
output rows the same plus "prompt" / "chosen" / "rejected" = the tokeniser's audio representation of each clip.
As in the reference, the 3*B clips of a batch go through the feature extractor as ONE padded batch in the order
prompts, chosens, rejecteds (pad_collate_fn, :50-54) -- the unit ids depend on batch composition, so the
order is part of the contract."""
import json
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from cli.extract_features import build_tokeniser  # noqa: E402
from slamkit_b200.audio_io import load_audio  # noqa: E402
from slamkit_b200.config import load_config, require  # noqa: E402

logger = logging.getLogger(__name__)
KEYS = ("prompt", "chosen", "rejected")


def read_triplets(path: str, skip=None, take=None):
    rows = [json.loads(line) for line in open(path) if line.strip()]
    if skip is not None:
        rows = rows[skip:]
    if take is not None:
        rows = rows[:take]
    return rows


def collate_triplets(rows, sample_rate: int = 16000, loader=load_audio):
    """pad_collate_fn: waveforms ordered [all prompts, all chosens, all rejecteds], right-padded with zeros."""
    wavs = [loader(r[f"{k}_path"], sample_rate) for k in KEYS for r in rows]
    lens = torch.tensor([len(w) for w in wavs])
    return torch.nn.utils.rnn.pad_sequence(wavs, batch_first=True, padding_value=0), lens


def attach(rows, reps):
    """Split the 3*B representations back into the rows (reference :74-81)."""
    n = len(rows)
    for i, r in enumerate(rows):
        for j, k in enumerate(KEYS):
            rep = reps[j * n + i]
            r[k] = {kk: (list(map(int, v)) if hasattr(v, "__len__") and not isinstance(v, str) else v) for kk, v in rep.items()} \
                if isinstance(rep, dict) else rep
    return rows


def main(argv=None):
    cfg = load_config("preference_alignment_feature_extractor", argv if argv is not None else sys.argv[1:])
    require(cfg, "data_path", "out_path")
    device = f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"
    rows = read_triplets(cfg.data_path, cfg.skip, cfg.take)
    tokeniser = build_tokeniser(cfg, device, max_batch=3 * cfg.batch_size)
    os.makedirs(os.path.dirname(os.path.abspath(cfg.out_path)), exist_ok=True)
    with open(cfg.out_path, "w") as f:
        for i in range(0, len(rows), cfg.batch_size):
            batch = rows[i:i + cfg.batch_size]
            wav, lens = collate_triplets(batch, cfg.sample_rate)
            for r in attach(batch, tokeniser.audio_represent(wav, lens)):
                f.write(json.dumps(r) + "\n")
    return cfg.out_path


if __name__ == "__main__":
    main()
