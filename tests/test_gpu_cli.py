"""End-to-end GPU test of the three drop-in entry points on synthetic data:
cli/extract_features.py (WAV dir -> features jsonl) -> cli/prepare_tokens.py -> cli/train.py (a few optimiser steps)."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_extract_prepare_train_pipeline(tmp_path):
    from cli import extract_features, prepare_tokens, train
    from slamkit_b200.audio_io import write_wav
    wav_dir = tmp_path / "audio"
    os.makedirs(wav_dir)
    g = torch.Generator().manual_seed(0)
    lens = [48000, 33000, 40000, 21000, 16000]
    for i, n in enumerate(lens):
        write_wav(str(wav_dir / f"utt{i}.wav"), (0.1 * torch.randn(n, generator=g)).clamp(-1, 1))
    feats = str(tmp_path / "features.jsonl")
    extract_features.main([f"data_path={wav_dir}", "ext=wav", f"out_path={feats}", "batch_size=2",
                           "tokeniser.feature_extractor_type=hubert_b200", "+synthetic_weights=true"])
    recs = [json.loads(l) for l in open(feats)]
    assert len(recs) == 5
    # descending-length order, file names preserved, durations sum to ceil(len/S * T) frames of each batch
    assert [os.path.basename(r["file_name"]) for r in recs] == ["utt0.wav", "utt2.wav", "utt1.wav", "utt3.wav", "utt4.wav"]
    for r in recs:
        assert len(r["units"]) == len(r["duration"]) > 0 and all(0 <= u < 500 for u in r["units"])
        assert all(a != b for a, b in zip(r["units"], r["units"][1:]))      # deduplicated
    assert sum(recs[0]["duration"]) == 75                                     # 48000 samples = 3 s -> 75 frames at 25 Hz
    out_dir = str(tmp_path / "tokens")
    tok_file = prepare_tokens.main([f"data_path={feats}", f"out_path={out_dir}"])
    lines = [json.loads(l) for l in open(tok_file)]
    assert len(lines) == 5 and lines[0]["audio_repr"].startswith("<Un")
    log = train.main([f"data.train_path={tok_file}", f"data.val_path={tok_file}", "model=slam", "model.tlm_type=b200",
                      "model.context_len=64", "model.config_args.twist_init=false",
                      "+model.shape.hidden=128", "+model.shape.n_layers=2", "+model.shape.n_heads=2",
                      "+model.shape.n_kv_heads=1", "+model.shape.ffn=256",
                      "training_args.per_device_train_batch_size=2", "+training_args.max_steps=12",
                      "+training_args.logging_steps=1", "training_args.warmup_steps=2", "training_args.warmup_ratio=0",
                      f"training_args.output_dir={tmp_path}/run"])
    losses = [r["loss"] for r in log if "loss" in r]
    assert len(losses) == 12 and losses[-1] < losses[0]
    assert os.path.exists(tmp_path / "run" / "model.safetensors") and os.path.exists(tmp_path / "run" / "config.json")
    st = json.load(open(tmp_path / "run" / "trainer_state.json"))
    assert st["global_step"] == 12 and st["eval_loss"] > 0 and st["num_input_tokens_seen"] > 0
    assert json.load(open(tmp_path / "run" / "tokeniser_config.json"))["num_units"] == 500
