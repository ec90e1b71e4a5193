"""Minimal audio file I/O for cli/extract_features.py: PCM WAV via the standard library (no torchaudio / soundfile in
the image).  Mirrors what `WavDataset.__getitem__` does after decoding (cli/extract_features.py:50-57): resample to the
target rate if needed, mix down to mono, return float32 in [-1, 1).  FLAC decoding is a SURVEY.md §8(f-3) item."""
from __future__ import annotations

import wave
from typing import Tuple

import numpy as np
import torch


def wav_num_frames(path: str) -> int:
    with wave.open(path, "rb") as w:
        return w.getnframes()


def load_wav(path: str, target_sr: int = 16000) -> torch.Tensor:
    with wave.open(path, "rb") as w:
        sr, ch, width, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        raw = w.readframes(n)
    if width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f"{path}: unsupported PCM sample width {width}")
    x = torch.from_numpy(x.reshape(-1, ch).T.copy())          # [channels, frames]
    if sr != target_sr:
        # band-limited resampling by linear interpolation of a sinc-free grid is NOT what torchaudio does; refuse rather
        # than silently produce different unit ids
        raise ValueError(f"{path}: sample rate {sr} != {target_sr}; resample offline (SURVEY.md §8 f-3)")
    return x.mean(dim=0) if x.shape[0] > 1 else x[0]


def write_wav(path: str, x: torch.Tensor, sr: int = 16000) -> None:
    pcm = (x.clamp(-1, 1 - 1 / 32768) * 32768.0).round().to(torch.int16).numpy()
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(pcm.tobytes())
