"""GPU parity of the causal-LM train step (hot path (ii)) through the handle-level C ABI, against oracle/lm_oracle.py
(pinned to the reference by tests/golden/lm_tiny.npz) and directly against the committed golden vectors.

Tolerances (floating point, bf16 compute): loss within 1e-3 relative (BASELINE.json north_star); logits and gradients
norm-wise within a small multiple of the bf16 rounding noise that two correct bf16 implementations show between each
other (the reference's own masked vs unmasked SDPA paths differ by 4.7e-3 on logits, see test_oracle_lm.py)."""
import os

import numpy as np
import pytest
import torch

from helpers import rel_err, u16_to_bf16

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _mk(cfg_o, seed, max_batch, max_seq):
    from slamkit_b200.lm import B200UnitLM, LMConfig
    from oracle import lm_oracle as O
    p = O.init_params(cfg_o, seed=seed)
    cfg = LMConfig(vocab_size=cfg_o.vocab_size, hidden=cfg_o.hidden, n_layers=cfg_o.n_layers, n_heads=cfg_o.n_heads,
                   n_kv_heads=cfg_o.n_kv_heads, head_dim=cfg_o.head_dim, ffn=cfg_o.ffn, rms_eps=cfg_o.rms_eps,
                   rope_theta=cfg_o.rope_theta, tie_embeddings=cfg_o.tie_embeddings, max_positions=2048)
    m = B200UnitLM(cfg, device=DEV, max_batch=max_batch, max_seq=max_seq)
    m.load_hf_state_dict(p)
    return m, p


def test_lm_matches_reference_golden(golden_dir):
    """Same weights / tokens as the fixture produced by the reference's UnitLM: loss, logits, grads, one optimiser step."""
    from oracle import lm_oracle as O
    from slamkit_b200.lm import B200AdamW
    z = np.load(os.path.join(golden_dir, "lm_tiny.npz"))
    c = z["cfg"]
    cfg_o = O.OracleLMConfig(vocab_size=int(c[0]), hidden=int(c[1]), n_layers=int(c[2]), n_heads=int(c[3]),
                             n_kv_heads=int(c[4]), head_dim=int(c[5]), ffn=int(c[6]))
    m, p = _mk(cfg_o, 123, 2, 48)
    ids, labels = torch.from_numpy(z["ids"]), torch.from_numpy(z["labels"])
    out = m.forward_backward(ids, labels, num_items_in_batch=float(z["num_items"]))
    loss = float(out.loss)
    assert abs(loss - float(z["loss"])) < 1e-3 * abs(float(z["loss"])), (loss, float(z["loss"]))
    valid = ids != 0
    logits = m.logits_view(2, 48).cpu()
    ref_logits = u16_to_bf16(z["logits_u16"])
    assert rel_err(logits[valid], ref_logits[valid]) < 8e-3
    sd_g = m.state_dict_hf(grads=True)
    worst = 0.0
    for k in p:
        ref = u16_to_bf16(z["grad::" + k]).view_as(p[k])
        worst = max(worst, rel_err(sd_g[k].cpu(), ref))
    assert worst < 2e-2, worst
    opt = B200AdamW(m, lr=1e-3, max_grad_norm=0.5)
    opt.step()
    assert abs(float(opt.stats[0]) - float(z["total_norm"])) < 0.01 * float(z["total_norm"])
    sd_p = m.state_dict_hf()
    for k in p:
        if k.endswith("k_proj.bias"):
            continue  # softmax is invariant to a key bias: its true gradient is 0, Adam turns rounding noise into +-lr
        ref = u16_to_bf16(z["new::" + k]).view_as(p[k])
        # the first AdamW step moves every weight by ~lr*sign(grad) = 1e-3, i.e. only ~8 bf16 ulps of a 0.02-sized
        # weight: compare the update direction and size, tolerant of sign flips on near-zero gradients
        upd, ref_upd = sd_p[k].cpu().float() - p[k].float(), ref.float() - p[k].float()
        assert rel_err(upd, ref_upd) < 0.2, k
        assert (sd_p[k].cpu() == ref).float().mean() > 0.9, k


@pytest.mark.parametrize("B,T,layers", [(2, 200, 3), (1, 1024, 2), (3, 130, 1)])
def test_lm_forward_backward_vs_oracle(B, T, layers):
    """Mid-size shapes (ragged T, GQA 4:2): loss / logits / every parameter gradient against the CPU oracle."""
    from oracle import lm_oracle as O
    cfg_o = O.OracleLMConfig(vocab_size=502, hidden=256, n_layers=layers, n_heads=4, n_kv_heads=2, head_dim=64, ffn=512)
    m, p = _mk(cfg_o, 5, B, T)
    g = torch.Generator().manual_seed(B * 1000 + T)
    ids = torch.randint(2, 502, (B, T), generator=g)
    ids[:, 0] = 1
    labels = ids.clone()
    if B > 1:
        ids[-1, T - 17:] = 0
        labels[-1, T - 17:] = -100
    n_items = float((labels != -100).sum())
    ref_loss, ref_logits, ref_grads = O.forward_backward(p, cfg_o, ids, labels, n_items)
    out = m.forward_backward(ids, labels, num_items_in_batch=n_items)
    assert abs(float(out.loss) - float(ref_loss)) < 1e-3 * abs(float(ref_loss)), (float(out.loss), float(ref_loss))
    assert int(out.stats[1]) == int((labels[:, 1:] != -100).sum())
    logits = m.logits_view(B, T).cpu()
    assert rel_err(logits, ref_logits) < 8e-3, rel_err(logits, ref_logits)
    sd_g = m.state_dict_hf(grads=True)
    errs = {k: rel_err(sd_g[k].cpu(), ref_grads[k]) for k in p}
    bad = {k: v for k, v in errs.items() if v > 2e-2 and not k.endswith("k_proj.bias")}  # d/d(k bias) == 0 exactly
    assert not bad, bad
    # gradient accumulation: a second identical micro-batch doubles the gradient
    m.forward_backward(ids, labels, num_items_in_batch=n_items, accumulate=True)
    sd_g2 = m.state_dict_hf(grads=True)
    for k in ("lm.model.layers.0.mlp.down_proj.weight", "lm.model.embed_tokens.weight", "lm.model.norm.weight",
              "lm.model.layers.0.self_attn.q_proj.bias"):
        assert rel_err(sd_g2[k].cpu().float(), 2 * sd_g[k].cpu().float()) < 8e-3, k


def test_lm_training_trajectory_vs_oracle():
    """Five optimiser steps (clip 0.5 + AdamW + cosine_with_min_lr): the loss trajectory follows the CPU oracle."""
    from oracle import lm_oracle as O
    from slamkit_b200.lm import B200AdamW, cosine_with_min_lr
    cfg_o = O.OracleLMConfig(vocab_size=502, hidden=128, n_layers=2, n_heads=2, n_kv_heads=1, head_dim=64, ffn=256)
    m, p = _mk(cfg_o, 9, 2, 64)
    tr = O.OracleTrainer(p, cfg_o, lr=1e-3, max_grad_norm=0.5)
    opt = B200AdamW(m, lr=1e-3, max_grad_norm=0.5)
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(2, 502, (2, 64), generator=g)
    ids[:, 0] = 1
    labels = ids.clone()
    n_items = float((labels != -100).sum())
    for step in range(5):
        lr = cosine_with_min_lr(step, base_lr=1e-3, min_lr=5e-5, warmup_steps=2, total_steps=10)
        lr = max(lr, 1e-4)
        ref = tr.train_step(ids, labels, lr=lr)
        out = m.forward_backward(ids, labels, num_items_in_batch=n_items)
        got = float(out.loss)
        opt.step(lr=lr)
        assert abs(got - ref) < 2e-3 * abs(ref), (step, got, ref)
    assert got < float(np.log(502)) - 0.05  # and it actually learns the repeated batch


def test_lm_full_size_properties():
    """BASELINE config-2 shape (Qwen2.5-0.5B body, [8,1024]): size-independent properties instead of an oracle run."""
    from slamkit_b200.lm import B200UnitLM, LMConfig, B200AdamW
    m = B200UnitLM(LMConfig(), device=DEV, max_batch=8, max_seq=1024, seed=0)
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(2, 502, (8, 1024), generator=g)
    ids[:, 0] = 1
    labels = ids.clone()
    out = m.forward_backward(ids, labels, num_items_in_batch=8192.0)
    loss0 = float(out.loss)
    # random-init model with tied embeddings (std 0.02, |h| = sqrt(896)): logits ~ N(0, 0.6^2), so the loss sits
    # ~sigma^2/2 above ln(502) (scaled by 8184/8192: 8 of the 8192 positions have no target)
    assert 6.2 < loss0 < 6.6, loss0
    assert int(out.stats[1]) == 8 * 1023
    g0 = m.grads.clone()
    assert torch.isfinite(g0.float()).all()
    # determinism: same batch -> bit-identical loss and gradients (no atomics on the GEMM/attention path)
    out2 = m.forward_backward(ids, labels, num_items_in_batch=8192.0)
    assert float(out2.loss) == loss0
    body = m.tensor("layers.0.wgu", grad=True).clone()
    m.forward_backward(ids, labels, num_items_in_batch=8192.0)
    assert torch.equal(body, m.tensor("layers.0.wgu", grad=True))
    # linearity of the backward pass in dloss
    m.forward_backward(ids, labels, num_items_in_batch=8192.0, loss_scale=2.0)
    assert rel_err(m.tensor("layers.5.wd", grad=True).float(), 2 * g0[m.tensors["layers.5.wd"][0]:][:896 * 4864].view(896, 4864).float()) < 5e-3
    # causality: changing the last token of every sequence leaves all earlier logits bit-identical
    lg1 = m.forward(ids).logits.clone()
    ids2 = ids.clone()
    ids2[:, -1] = (ids2[:, -1] + 7) % 500 + 2
    lg2 = m.forward(ids2).logits
    assert torch.equal(lg1[:, :-1], lg2[:, :-1]) and not torch.equal(lg1[:, -1], lg2[:, -1])
    # batch independence: sequence 3 alone gives the same logits as inside the batch
    lg3 = m.forward(ids[3:4]).logits
    assert rel_err(lg3[0].float(), lg1[3].float()) < 1e-6
    # a few optimiser steps on the same batch reduce the loss
    opt = B200AdamW(m, lr=1e-3, max_grad_norm=0.5)
    for _ in range(3):
        m.forward_backward(ids, labels, num_items_in_batch=8192.0)
        opt.step()
    assert float(m.forward_backward(ids, labels, num_items_in_batch=8192.0).loss) < loss0 - 0.05


def test_lm_packed_batch_equals_separate_documents():
    """Packing (DataCollatorWithFlattening: one row, position_ids restarting per document): logits and loss of the
    packed row equal those of the documents run one by one, and match the oracle's block-diagonal restatement."""
    from oracle import lm_oracle as O
    cfg_o = O.OracleLMConfig(vocab_size=502, hidden=128, n_layers=2, n_heads=2, n_kv_heads=1, head_dim=64, ffn=256)
    m, p = _mk(cfg_o, 5, 2, 512)
    g = torch.Generator().manual_seed(3)
    lens = [37, 128, 1, 200, 90]
    docs = [torch.randint(2, 502, (n,), generator=g) for n in lens]
    ids = torch.cat(docs)[None]                                             # [1, 456]
    pos = torch.cat([torch.arange(n) for n in lens])[None]
    labels = ids.clone()
    for a in np.cumsum([0] + lens[:-1]):
        labels[0, a] = -100                                                 # separator: first token of each document
    n_items = float((labels[:, 1:] != -100).sum())
    out = m.forward_backward(ids, labels, position_ids=pos, num_items_in_batch=n_items)
    loss_packed = float(out.loss)
    grads_packed = m.grads.clone()
    lg_packed = m.forward(ids, position_ids=pos).logits[0].float().cpu()
    # (1) the documents one by one (plain causal kernels)
    off, nll = 0, 0.0
    for n, d in zip(lens, docs):
        if n > 1:
            lg = m.forward(d[None]).logits[0].float().cpu()
            assert rel_err(lg_packed[off:off + n], lg) < 8e-3, (n, rel_err(lg_packed[off:off + n], lg))
            nll += float(torch.nn.functional.cross_entropy(lg[:-1], d[1:], reduction="sum"))
        off += n
    assert abs(loss_packed - nll / n_items) < 2e-3 * abs(nll / n_items), (loss_packed, nll / n_items)
    # (2) the oracle with the explicit block-diagonal mask
    lo, _, go = O.forward_backward(p, cfg_o, ids, labels, n_items, position_ids=pos, packed=True)
    assert abs(loss_packed - float(lo)) < 1e-3 * abs(float(lo)), (loss_packed, float(lo))
    got = m.state_dict_hf(grads=True) if hasattr(m, "state_dict_hf") else None
    if got is not None:
        for k in ("lm.model.layers.0.mlp.down_proj.weight", "lm.model.layers.1.self_attn.q_proj.weight",
                  "lm.model.layers.0.self_attn.v_proj.weight"):
            assert rel_err(got[k].float().cpu(), go[k].float()) < 3e-2, (k, rel_err(got[k].float().cpu(), go[k].float()))
    # (3) without position_ids the same row is one long document: different logits after the first boundary
    lg_plain = m.forward(ids).logits[0].float().cpu()
    assert rel_err(lg_plain[:lens[0]], lg_packed[:lens[0]]) < 1e-6 and rel_err(lg_plain[lens[0]:], lg_packed[lens[0]:]) > 1e-2
    assert torch.isfinite(grads_packed.float()).all()


def test_lm_packed_matches_reference_golden(golden_dir):
    """tests/golden/lm_packed.npz (reference UnitLM with the explicit block-diagonal mask): loss and logits of the packed
    row through sk_lm_forward with position_ids."""
    from oracle import lm_oracle as O
    z = np.load(os.path.join(golden_dir, "lm_packed.npz"))
    c = z["cfg"]
    cfg_o = O.OracleLMConfig(vocab_size=int(c[0]), hidden=int(c[1]), n_layers=int(c[2]), n_heads=int(c[3]),
                             n_kv_heads=int(c[4]), head_dim=int(c[5]), ffn=int(c[6]))
    m, _ = _mk(cfg_o, 123, 1, 128)
    ids, pos, labels = (torch.from_numpy(z[k]) for k in ("ids", "position_ids", "labels"))
    out = m.forward(ids, position_ids=pos, labels=labels, num_items_in_batch=float(z["num_items"]))
    assert abs(float(out.loss) - float(z["loss"])) < 1e-3 * abs(float(z["loss"])), (float(out.loss), float(z["loss"]))
    assert rel_err(out.logits[0].float().cpu(), u16_to_bf16(z["logits_u16"])[0].float()) < 8e-3


@pytest.mark.parametrize("head_chunk", [0, 128])
def test_lm_large_vocabulary_vs_oracle(head_chunk, monkeypatch):
    """A vocabulary far above the unit-only 502 (the interleaved text+unit configuration): lm_head GEMMs with thousands
    of columns, the block-per-row CE kernel, a larger tied embedding in the optimiser.  head_chunk = 128 forces the
    chunked lm_head + CE that 152 k-column vocabularies use by default (logits exist one 128-row chunk at a time, the
    gradient is written over them, dE accumulates over the chunks): 192 rows = one full and one ragged chunk."""
    from oracle import lm_oracle as O
    monkeypatch.setenv("SK_HEAD_CHUNK", str(head_chunk))
    cfg_o = O.OracleLMConfig(vocab_size=4099, hidden=128, n_layers=2, n_heads=2, n_kv_heads=1, head_dim=64, ffn=256)
    m, p = _mk(cfg_o, 7, 2, 96)
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(1, 4099, (2, 96), generator=g)
    labels = ids.clone()
    labels[1, 70:] = -100
    n_items = float((labels[:, 1:] != -100).sum())
    out = m.forward_backward(ids, labels, num_items_in_batch=n_items)
    lo, lg_o, go = O.forward_backward(p, cfg_o, ids, labels, n_items)
    assert abs(float(out.loss) - float(lo)) < 1e-3 * abs(float(lo)), (float(out.loss), float(lo))
    lg = m.forward(ids).logits.float().cpu()
    assert lg.shape[-1] == 4099 and rel_err(lg, lg_o.float()) < 8e-3
    got = m.state_dict_hf(grads=True)
    for k in ("lm.model.embed_tokens.weight", "lm.model.layers.1.mlp.down_proj.weight", "lm.model.norm.weight"):
        assert rel_err(got[k].float().cpu(), go[k].float()) < 3e-2, (k, rel_err(got[k].float().cpu(), go[k].float()))


def test_grad_norm_and_clip_matches_torch():
    from oracle import lm_oracle as O
    from slamkit_b200.lm import B200AdamW
    cfg_o = O.OracleLMConfig(vocab_size=502, hidden=128, n_layers=2, n_heads=2, n_kv_heads=1, head_dim=64, ffn=256)
    m, p = _mk(cfg_o, 1, 1, 64)
    gen = torch.Generator().manual_seed(0)
    fake = {k: (torch.randn(v.shape, generator=gen) * 0.01).to(torch.bfloat16) for k, v in p.items()}
    m.grads.zero_()
    m.load_hf_state_dict(fake, grads=True)
    opt = B200AdamW(m, lr=0.0, max_grad_norm=0.5)
    opt.step()
    total = O.clip_grad_norm_([v.clone() for v in fake.values()], 0.5)
    assert float(opt.stats[0]) == float(total), (float(opt.stats[0]), float(total))  # bf16-emulated total norm, exact
    exact = torch.sqrt(sum((v.float() ** 2).sum() for v in fake.values()))
    assert abs(float(opt.stats[2]) - float(exact)) < 1e-4 * float(exact)


def test_hf_layout_checkpoint_roundtrip(tmp_path):
    """save_pretrained writes the reference's UnitLM layout (lm.-prefixed safetensors + UnitLMConfig json): reloading
    gives bit-identical logits, and the tensor names / shapes are exactly those of the oracle's HF-named parameters."""
    import json
    from safetensors.torch import load_file
    from oracle import lm_oracle as O
    from slamkit_b200.lm import B200UnitLM
    cfg_o = O.OracleLMConfig(vocab_size=502, hidden=128, n_layers=2, n_heads=2, n_kv_heads=1, head_dim=64, ffn=256)
    m, p = _mk(cfg_o, 3, 2, 64)
    d = str(tmp_path / "ckpt")
    m.save_pretrained(d)
    sd = load_file(d + "/model.safetensors")
    assert set(sd.keys()) == set(p.keys())
    for k in p:
        assert sd[k].shape == p[k].shape and torch.equal(sd[k], p[k]), k
    c = json.load(open(d + "/config.json"))
    assert c["model_type"] == "speech_language_model" and c["base_config"]["rope_parameters"]["rope_theta"] == 10000.0
    m2 = B200UnitLM.from_pretrained(d, device=DEV, max_batch=2, max_seq=64)
    ids = torch.randint(2, 502, (2, 64), generator=torch.Generator().manual_seed(0))
    assert torch.equal(m.forward(ids).logits.clone(), m2.forward(ids).logits)
