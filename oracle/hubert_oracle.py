"""CPU oracle for hot path (i): waveform -> HuBERT-25Hz layer-`layer` features -> k-means unit ids -> dedup.

TEST INFRASTRUCTURE ONLY.  Nothing under slamkit_b200/ may import this module; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / `--impl reference` legs use it, and only as the checker or the timed CPU baseline.

Restates, stage by stage in fp32 torch / numpy, what `HubertFeatureExtractor.extract`
(slamkit/feature_extractor/hubert_feature_extractor.py:40-50) executes inside third-party code that is not vendored
under /root/reference: HF `HubertModel` (transformers >=4.48.1 per pyproject.toml:13; 5.5.0 installed;
HF:models/hubert/modeling_hubert.py:45-231,262-470) and `sklearn.cluster.KMeans.predict` (scikit-learn unpinned; 1.9.0
installed; SK:cluster/_kmeans.py:1075-1107, SK:cluster/_k_means_lloyd.pyx:168-213), then the run-length dedup of
`UnitTokeniser.audio_represent` (slamkit/tokeniser/unit_tokeniser.py:54-60).

Pinned by tests/golden/hubert_tiny.npz (oracle/make_goldens.py: the reference's own extract() on seeded weights) and by
the reference's golden files example_data/{features,tokens}.jsonl for the frame-count / rel_l / dedup / string rules.
Real-weight unit ids (mHuBERT-25Hz + km500 checkpoints) cannot be pinned offline: "parity unpinned" for those.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from itertools import groupby
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class OracleHubertConfig:
    """mHuBERT-25Hz geometry by default (config/tokeniser/feature_extractor/mhubert_25.yaml; conv stack pinned by the
    352/398-frame known answer of example_data/features.jsonl)."""
    conv_dim: int = 512
    conv_kernel: Tuple[int, ...] = (10, 3, 3, 3, 3, 2, 2, 2)
    conv_stride: Tuple[int, ...] = (5, 2, 2, 2, 2, 2, 2, 2)
    hidden: int = 768
    n_heads: int = 12
    ffn: int = 3072
    n_layers: int = 12
    pos_conv_kernel: int = 128
    pos_conv_groups: int = 16
    n_units: int = 500
    layer: int = 11
    ln_eps: float = 1e-5
    pad: int = 40   # F.pad(wav, (40, 40)), hubert_feature_extractor.py:42


def frame_counts(cfg: OracleHubertConfig, n_samples: int) -> List[int]:
    """Conv1d output lengths per layer for a (pad,pad)-padded clip (HF:modeling_hubert.py `_conv_out_length`)."""
    L = n_samples + 2 * cfg.pad
    out = []
    for k, s in zip(cfg.conv_kernel, cfg.conv_stride):
        L = (L - k) // s + 1
        out.append(L)
    return out


def rel_lengths(lens: torch.Tensor, n_samples: int, n_frames: int) -> torch.Tensor:
    """hubert_feature_extractor.py:46 -- float32 arithmetic, then ceil().int()."""
    return ((lens.float() / n_samples) * n_frames).ceil().int()


def init_hubert_params(cfg: OracleHubertConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(shape, generator=g) * std

    C, H, Fd = cfg.conv_dim, cfg.hidden, cfg.ffn
    p: Dict[str, torch.Tensor] = {}
    p["conv0.weight"] = rn(C, 1, cfg.conv_kernel[0], std=math.sqrt(2.0 / cfg.conv_kernel[0]))
    p["gn.weight"] = 1.0 + 0.1 * rn(C)
    p["gn.bias"] = 0.1 * rn(C)
    for i in range(1, len(cfg.conv_kernel)):
        p[f"conv{i}.weight"] = rn(C, C, cfg.conv_kernel[i], std=math.sqrt(2.0 / (C * cfg.conv_kernel[i])))
    p["fp.ln.weight"], p["fp.ln.bias"] = 1.0 + 0.1 * rn(C), 0.1 * rn(C)
    p["fp.proj.weight"], p["fp.proj.bias"] = rn(H, C, std=1.0 / math.sqrt(C)), 0.02 * rn(H)
    K, G = cfg.pos_conv_kernel, cfg.pos_conv_groups
    p["pos.v"] = rn(H, H // G, K, std=math.sqrt(4.0 / (K * H)))
    p["pos.g"] = p["pos.v"].norm(dim=(0, 1), keepdim=True) * (1.0 + 0.1 * rn(1, 1, K))
    p["pos.bias"] = 0.02 * rn(H)
    p["enc.ln.weight"], p["enc.ln.bias"] = 1.0 + 0.1 * rn(H), 0.1 * rn(H)
    for l in range(cfg.n_layers):
        q = f"layers.{l}."
        for nm in ("q", "k", "v", "o"):
            p[q + nm + ".weight"], p[q + nm + ".bias"] = rn(H, H, std=1.0 / math.sqrt(H)), 0.02 * rn(H)
        p[q + "ln1.weight"], p[q + "ln1.bias"] = 1.0 + 0.1 * rn(H), 0.1 * rn(H)
        p[q + "ff1.weight"], p[q + "ff1.bias"] = rn(Fd, H, std=1.0 / math.sqrt(H)), 0.02 * rn(Fd)
        p[q + "ff2.weight"], p[q + "ff2.bias"] = rn(H, Fd, std=1.0 / math.sqrt(Fd)), 0.02 * rn(H)
        p[q + "ln2.weight"], p[q + "ln2.bias"] = 1.0 + 0.1 * rn(H), 0.1 * rn(H)
    p["kmeans.centers"] = rn(cfg.n_units, H)
    return p


def hf_state_dict_from_oracle(p: Dict[str, torch.Tensor], cfg: OracleHubertConfig, template: Dict[str, torch.Tensor]):
    """Oracle parameter names -> HF HubertModel.state_dict() names (used only by oracle/make_goldens.py)."""
    sd = {}
    sd["feature_extractor.conv_layers.0.conv.weight"] = p["conv0.weight"]
    sd["feature_extractor.conv_layers.0.layer_norm.weight"] = p["gn.weight"]
    sd["feature_extractor.conv_layers.0.layer_norm.bias"] = p["gn.bias"]
    for i in range(1, len(cfg.conv_kernel)):
        sd[f"feature_extractor.conv_layers.{i}.conv.weight"] = p[f"conv{i}.weight"]
    sd["feature_projection.layer_norm.weight"], sd["feature_projection.layer_norm.bias"] = p["fp.ln.weight"], p["fp.ln.bias"]
    sd["feature_projection.projection.weight"], sd["feature_projection.projection.bias"] = p["fp.proj.weight"], p["fp.proj.bias"]
    sd["encoder.pos_conv_embed.conv.bias"] = p["pos.bias"]
    sd["encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = p["pos.g"]
    sd["encoder.pos_conv_embed.conv.parametrizations.weight.original1"] = p["pos.v"]
    sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"] = p["enc.ln.weight"], p["enc.ln.bias"]
    for l in range(cfg.n_layers):
        q, h = f"layers.{l}.", f"encoder.layers.{l}."
        for a, b in (("q", "attention.q_proj"), ("k", "attention.k_proj"), ("v", "attention.v_proj"),
                     ("o", "attention.out_proj"), ("ff1", "feed_forward.intermediate_dense"),
                     ("ff2", "feed_forward.output_dense"), ("ln1", "layer_norm"), ("ln2", "final_layer_norm")):
            sd[h + b + ".weight"], sd[h + b + ".bias"] = p[q + a + ".weight"], p[q + a + ".bias"]
    for k in template:
        if k not in sd:
            sd[k] = template[k]   # e.g. masked_spec_embed (unused at inference)
    return sd


def conv_feature_encoder(p, cfg: OracleHubertConfig, wav_padded: torch.Tensor) -> torch.Tensor:
    """HubertFeatureEncoder with feat_extract_norm='group' (HF:modeling_hubert.py:106-125,154-213): conv0 -> GroupNorm
    (one group per channel = statistics over TIME, zero-padded tail included) -> GELU; then 7x [conv, GELU].
    [B,S] -> [B,C,T]."""
    h = wav_padded[:, None]
    h = F.conv1d(h, p["conv0.weight"], stride=cfg.conv_stride[0])
    h = F.group_norm(h, cfg.conv_dim, p["gn.weight"], p["gn.bias"], eps=1e-5)
    h = F.gelu(h)
    for i in range(1, len(cfg.conv_kernel)):
        h = F.gelu(F.conv1d(h, p[f"conv{i}.weight"], stride=cfg.conv_stride[i]))
    return h


def pos_conv_weight(p) -> torch.Tensor:
    """torch weight_norm(dim=2): w = g * v / ||v||, norm over every dim except 2."""
    v = p["pos.v"]
    return p["pos.g"] * v / v.norm(dim=(0, 1), keepdim=True)


def encoder_embed(p, cfg: OracleHubertConfig, feats_bct: torch.Tensor) -> torch.Tensor:
    """transpose -> HubertFeatureProjection (LN + Linear, HF:216-231) -> x + GELU(pos_conv(x)) (HF:45-92, the last frame
    of the even-kernel conv output dropped) -> encoder LayerNorm (HF:440-443).  Returns hidden_states[0]: [B,T,H]."""
    x = feats_bct.transpose(1, 2)
    x = F.layer_norm(x, (cfg.conv_dim,), p["fp.ln.weight"], p["fp.ln.bias"], cfg.ln_eps)
    x = F.linear(x, p["fp.proj.weight"], p["fp.proj.bias"])
    pc = F.conv1d(x.transpose(1, 2), pos_conv_weight(p), p["pos.bias"], padding=cfg.pos_conv_kernel // 2,
                  groups=cfg.pos_conv_groups)
    if cfg.pos_conv_kernel % 2 == 0:
        pc = pc[:, :, :-1]
    x = x + F.gelu(pc).transpose(1, 2)
    return F.layer_norm(x, (cfg.hidden,), p["enc.ln.weight"], p["enc.ln.bias"], cfg.ln_eps)


def encoder_layer(p, cfg: OracleHubertConfig, l: int, x: torch.Tensor) -> torch.Tensor:
    """HubertEncoderLayer (post-LN, HF:372-405), bidirectional attention with NO mask (the reference passes none)."""
    q_ = f"layers.{l}."
    B, T, H = x.shape
    hd = H // cfg.n_heads

    def heads(t):
        return t.view(B, T, cfg.n_heads, hd).transpose(1, 2)

    q = heads(F.linear(x, p[q_ + "q.weight"], p[q_ + "q.bias"]))
    k = heads(F.linear(x, p[q_ + "k.weight"], p[q_ + "k.bias"]))
    v = heads(F.linear(x, p[q_ + "v.weight"], p[q_ + "v.bias"]))
    a = F.scaled_dot_product_attention(q, k, v, scale=hd ** -0.5).transpose(1, 2).reshape(B, T, H)
    x = x + F.linear(a, p[q_ + "o.weight"], p[q_ + "o.bias"])
    x = F.layer_norm(x, (H,), p[q_ + "ln1.weight"], p[q_ + "ln1.bias"], cfg.ln_eps)
    f = F.linear(F.gelu(F.linear(x, p[q_ + "ff1.weight"], p[q_ + "ff1.bias"])), p[q_ + "ff2.weight"], p[q_ + "ff2.bias"])
    return F.layer_norm(x + f, (H,), p[q_ + "ln2.weight"], p[q_ + "ln2.bias"], cfg.ln_eps)


@torch.inference_mode()
def features(p, cfg: OracleHubertConfig, wav: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """hidden_states[cfg.layer] of HubertModel(F.pad(wav,(pad,pad)), output_hidden_states=True): only `layer` encoder
    layers are needed (hidden_states[i] = output of layer i-1). [B,S] fp32 -> [B,T,H] fp32."""
    x = conv_feature_encoder(p, cfg, F.pad(wav, (cfg.pad, cfg.pad)))
    if taps is not None:
        taps["conv"] = x
    x = encoder_embed(p, cfg, x)
    if taps is not None:
        taps["embed"] = x
    for l in range(cfg.layer):
        x = encoder_layer(p, cfg, l, x)
        if taps is not None:
            taps[f"layer{l}"] = x
    return x


def kmeans_predict(x: np.ndarray, centers: np.ndarray, chunk: int = 256) -> np.ndarray:
    """sklearn KMeans.predict on dense fp32 data (SK:cluster/_k_means_lloyd.pyx:168-213 `_update_chunk_dense`):
    per 256-row chunk, pairwise = ||C_j||^2 - 2 x.C_j (the ||x||^2 term is dropped), label = first j with the strictly
    smallest value.  Everything in the dtype of x (fp32)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    c = np.ascontiguousarray(centers, dtype=np.float32)
    csq = np.einsum("ij,ij->i", c, c).astype(np.float32)
    out = np.empty(x.shape[0], dtype=np.int32)
    for s in range(0, x.shape[0], chunk):
        d = csq[None, :] + np.float32(-2.0) * (x[s:s + chunk] @ c.T)
        out[s:s + chunk] = np.argmin(d, axis=1)  # np.argmin returns the first minimum, like the strict '<' scan
    return out


def kmeans_margins(x: np.ndarray, centers: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """(labels, top-2 margin of the squared-distance surrogate) in float64: used to classify id mismatches as ties."""
    x = x.astype(np.float64)
    c = centers.astype(np.float64)
    d = (c * c).sum(1)[None, :] - 2.0 * x @ c.T
    part = np.partition(d, 1, axis=1)
    return np.argmin(d, axis=1).astype(np.int32), part[:, 1] - part[:, 0]


def extract(p, cfg: OracleHubertConfig, wav: torch.Tensor, lens: Optional[torch.Tensor] = None) -> List[np.ndarray]:
    """HubertFeatureExtractor.extract (hubert_feature_extractor.py:40-50) + batch_cluster (:73-81)."""
    cont = features(p, cfg, wav).numpy()
    B, T, C = cont.shape
    toks = kmeans_predict(cont.reshape(B * T, C), p["kmeans.centers"].numpy()).reshape(B, T)
    if lens is not None:
        rel = rel_lengths(lens, wav.shape[1], T)
    else:
        rel = [T] * B
    return [t[:int(l)] for t, l in zip(toks, rel)]


def dedup(units: Sequence[int]) -> Tuple[List[int], List[int]]:
    """UnitTokeniser.audio_represent with dedup=True (unit_tokeniser.py:57): run-length encode."""
    u, d = [], []
    for k, g in groupby(list(units)):
        u.append(int(k))
        d.append(len(list(g)))
    return u, d


def stringify(units: Sequence[int]) -> str:
    """UnitTokeniser.stringify_representation (unit_tokeniser.py:62-63)."""
    return "".join(f"<Un{u}>" for u in units)


def token_ids(units: Sequence[int], bos_eos: int = 1, pad: int = 0) -> List[int]:
    """WordLevel vocab of UnitTokeniser._init_text_tokeniser (unit_tokeniser.py:33-47): <Un{i}> -> i + offset with
    offset = max(eos,bos,pad)+1; template '<S> $0 <S>'."""
    off = max(bos_eos, pad) + 1
    return [bos_eos] + [int(u) + off for u in units] + [bos_eos]
