"""Pins oracle/lm_oracle.py against tests/golden/lm_tiny.npz, which oracle/make_goldens.py produced by running the
reference's own `slamkit.model.unit_lm.UnitLM` (+ HF Qwen2, torch AdamW) in the build container. CPU only."""
import os

import numpy as np
import torch

from oracle import lm_oracle as O
from helpers import rel_err, u16_to_bf16


def _load(golden_dir):
    z = np.load(os.path.join(golden_dir, "lm_tiny.npz"))
    c = z["cfg"]
    cfg = O.OracleLMConfig(vocab_size=int(c[0]), hidden=int(c[1]), n_layers=int(c[2]), n_heads=int(c[3]),
                           n_kv_heads=int(c[4]), head_dim=int(c[5]), ffn=int(c[6]))
    return z, cfg


def test_forward_loss_logits_match_reference(golden_dir):
    z, cfg = _load(golden_dir)
    p = O.init_params(cfg, seed=123)
    ids, labels = torch.from_numpy(z["ids"]), torch.from_numpy(z["labels"])
    logits = O.forward_logits(p, cfg, ids)
    # Without an attention_mask the reference takes the same is_causal SDPA path: bit-identical bf16 logits.
    assert torch.equal(logits, u16_to_bf16(z["logits_nomask_u16"]))
    # With the collator's attention_mask (the Trainer path) HF builds an explicit mask: non-pad rows agree to bf16
    # rounding; right-padded query rows additionally mask the pad KEYS (they carry label -100 and feed nothing).
    ref_logits = u16_to_bf16(z["logits_u16"])
    valid = ids != 0
    assert rel_err(logits[valid], ref_logits[valid]) < 4e-3
    loss = O.compute_loss(logits, labels, float(z["num_items"]))
    assert abs(float(loss) - float(z["loss"])) <= 1e-6 * abs(float(z["loss"]))


def test_backward_grads_match_reference(golden_dir):
    z, cfg = _load(golden_dir)
    p = O.init_params(cfg, seed=123)
    ids, labels = torch.from_numpy(z["ids"]), torch.from_numpy(z["labels"])
    _, _, grads = O.forward_backward(p, cfg, ids, labels, float(z["num_items"]))
    for k, g in grads.items():
        ref = u16_to_bf16(z["grad::" + k]).view_as(g)
        assert rel_err(g, ref) < 2e-3, k  # autocast vs plain bf16 autograd may differ by bf16 rounding order only


def test_optimizer_step_matches_reference(golden_dir):
    z, cfg = _load(golden_dir)
    p = O.init_params(cfg, seed=123)
    ids, labels = torch.from_numpy(z["ids"]), torch.from_numpy(z["labels"])
    tr = O.OracleTrainer(p, cfg, lr=1e-3, max_grad_norm=0.5)
    loss = tr.train_step(ids, labels)
    assert abs(loss - float(z["loss"])) <= 1e-6 * abs(loss)
    assert abs(float(tr.last_total_norm) - float(z["total_norm"])) <= 0.01 * float(z["total_norm"])
    for k, v in tr.p.items():
        ref = u16_to_bf16(z["new::" + k]).view_as(v)
        # first AdamW step moves every weight by ~lr; compare the UPDATE, not the weight
        upd, ref_upd = v.float() - p[k].float(), ref.float() - p[k].float()
        assert rel_err(upd, ref_upd) < 0.02, k


def test_compute_loss_mean_and_ignore_index():
    torch.manual_seed(0)
    logits = torch.randn(2, 5, 11).bfloat16()
    labels = torch.randint(0, 11, (2, 5))
    labels[0, 3:] = -100
    a = O.compute_loss(logits, labels)  # mean over valid shifted targets
    sl, st = logits.float()[:, :-1].reshape(-1, 11), labels[:, 1:].reshape(-1)
    keep = st != -100
    b = -(torch.log_softmax(sl, -1)[keep, st[keep]]).mean()
    assert abs(float(a) - float(b)) < 1e-6
    n = float((labels != -100).sum())
    c = O.compute_loss(logits, labels, n)
    assert abs(float(c) - float(b) * keep.sum().item() / n) < 1e-6


def test_cosine_with_min_lr_matches_hf():
    from transformers.optimization import get_cosine_with_min_lr_schedule_with_warmup

    w = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([w], lr=1e-3)
    sch = get_cosine_with_min_lr_schedule_with_warmup(opt, 100, 1000, min_lr=5e-5)
    for s in range(0, 1000, 37):
        want = 1e-3 * sch.lr_lambdas[0](s)
        got = O.cosine_with_min_lr(s, base_lr=1e-3, min_lr=5e-5, warmup_steps=100, total_steps=1000)
        assert abs(want - got) < 1e-12


def test_packed_mask_equals_running_documents_separately():
    """The oracle's restatement of the reference's packing path (DataCollatorWithFlattening + varlen flash attention,
    slamkit/data/hf_dataset.py:61-62): a packed row with per-document position_ids gives each document exactly the
    logits it gets on its own (fp32 so that the comparison is not blurred by bf16 summation order)."""
    cfg = O.OracleLMConfig(vocab_size=64, hidden=64, n_layers=2, n_heads=2, n_kv_heads=1, head_dim=32, ffn=128)
    p = {k: v.float() for k, v in O.init_params(cfg, seed=0).items()}
    g = torch.Generator().manual_seed(0)
    lens = [5, 1, 17, 9]
    docs = [torch.randint(2, 64, (n,), generator=g) for n in lens]
    ids = torch.cat(docs)[None]
    pos = torch.cat([torch.arange(n) for n in lens])[None]
    doc = O.document_ids(pos.clone())
    assert doc.tolist() == [sum(([i + 1] * n for i, n in enumerate(lens)), [])]
    packed = O.forward_logits(p, cfg, ids, pos, packed=True)[0]
    off = 0
    for n, d in zip(lens, docs):
        alone = O.forward_logits(p, cfg, d[None])[0]
        assert float((packed[off:off + n] - alone).abs().max()) < 1e-4
        off += n
    leaky = O.forward_logits(p, cfg, ids, pos, packed=False)[0]       # same row without the document mask
    assert float((leaky[lens[0]:] - packed[lens[0]:]).abs().max()) > 1e-3


def test_packed_batch_matches_reference_golden(golden_dir):
    """tests/golden/lm_packed.npz: the reference's UnitLM (HF Qwen2) on a packed row with the explicit block-diagonal
    causal 4-D mask -- the oracle's `packed=True` path reproduces its logits bit for bit and its loss to 1e-6."""
    z = np.load(os.path.join(golden_dir, "lm_packed.npz"))
    c = z["cfg"]
    cfg = O.OracleLMConfig(vocab_size=int(c[0]), hidden=int(c[1]), n_layers=int(c[2]), n_heads=int(c[3]),
                           n_kv_heads=int(c[4]), head_dim=int(c[5]), ffn=int(c[6]))
    p = O.init_params(cfg, seed=123)
    ids, pos, labels = (torch.from_numpy(z[k]) for k in ("ids", "position_ids", "labels"))
    logits = O.forward_logits(p, cfg, ids, pos, packed=True)
    assert torch.equal(logits, u16_to_bf16(z["logits_u16"]))
    loss = O.compute_loss(logits, labels, float(z["num_items"]))
    assert abs(float(loss) - float(z["loss"])) < 1e-6 * abs(float(z["loss"]))
    assert O.document_ids(pos.clone())[0, -1] == len(z["lens"])
