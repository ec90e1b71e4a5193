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


# ---- FLAC through the library's host-side decoder (sk_flac_*) ---------------------------------------------------------
def flac_info(path: str):
    import ctypes as C
    from . import _lib as L
    lib = L.load()
    sr, ch, bps, n = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
    md5 = (C.c_uint8 * 16)()
    L.check(lib.sk_flac_info(path.encode(), C.byref(sr), C.byref(ch), C.byref(bps), C.byref(n), md5))
    return {"sample_rate": sr.value, "channels": ch.value, "bits_per_sample": bps.value, "num_frames": n.value,
            "md5": bytes(md5)}


def flac_decode_int(path: str) -> np.ndarray:
    """Interleaved PCM as int32 [frames, channels]."""
    import ctypes as C
    from . import _lib as L
    lib = L.load()
    info = flac_info(path)
    cap = max(info["num_frames"], 1) + 65536
    buf = np.empty((cap, info["channels"]), dtype=np.int32)
    n = C.c_int64()
    L.check(lib.sk_flac_decode_i32(path.encode(), buf.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int64(cap), C.byref(n)))
    return buf[:n.value]


def load_flac(path: str, target_sr: int = 16000) -> torch.Tensor:
    """torchaudio.load semantics for integer FLAC: float32 = int / 2^(bits-1), then the channel mean."""
    info = flac_info(path)
    if info["sample_rate"] != target_sr:
        raise ValueError(f"{path}: sample rate {info['sample_rate']} != {target_sr}; resample offline")
    pcm = flac_decode_int(path).astype(np.float32) / float(1 << (info["bits_per_sample"] - 1))
    x = torch.from_numpy(pcm.T.copy())
    return x.mean(dim=0) if x.shape[0] > 1 else x[0]


def load_audio(path: str, target_sr: int = 16000) -> torch.Tensor:
    return load_flac(path, target_sr) if path.lower().endswith(".flac") else load_wav(path, target_sr)


def audio_num_frames(path: str) -> int:
    return flac_info(path)["num_frames"] if path.lower().endswith(".flac") else wav_num_frames(path)
