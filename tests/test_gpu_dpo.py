"""DPO step on the GPU path against autograd of the CPU restatement (oracle/dpo_oracle.py)."""
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_tokenize_row_and_collate():
    from slamkit_b200.dpo import collate_pairs, tokenize_row
    from slamkit_b200.tokeniser import B200UnitTokeniser
    tok = B200UnitTokeniser(None, load_fe=False)
    row = tokenize_row({"prompt": "<Un3><Un4><Un5>", "chosen": "<Un7><Un8>", "rejected": "<Un9>"}, tok, 2, 2)
    assert row == {"prompt_input_ids": [6, 7], "chosen_input_ids": [9, 10], "rejected_input_ids": [11, 1]}
    row = tokenize_row({"prompt": "<Un3>", "chosen": "<Un7>", "rejected": "<Un9><Un1>"}, tok, None, None)
    assert row == {"prompt_input_ids": [1, 5], "chosen_input_ids": [9, 1], "rejected_input_ids": [11, 3, 1]}
    ids, labels = collate_pairs([row])
    assert ids.tolist() == [[1, 5, 9, 1, 0], [1, 5, 11, 3, 1]]
    assert labels.tolist() == [[-100, -100, 9, 1, -100], [-100, -100, 11, 3, 1]]


def test_dpo_step_matches_oracle_autograd():
    from oracle import dpo_oracle as D
    from oracle import lm_oracle as O
    from slamkit_b200.dpo import B200DPOTrainer
    from slamkit_b200.lm import B200UnitLM, LMConfig
    cfg_o = O.OracleLMConfig(vocab_size=502, hidden=128, n_layers=2, n_heads=2, n_kv_heads=1, head_dim=64, ffn=256)
    ref_p = O.init_params(cfg_o, seed=1)
    g = torch.Generator().manual_seed(2)
    pol_p = {k: (v.float() + 0.01 * torch.randn(v.shape, generator=g)).to(torch.bfloat16) for k, v in ref_p.items()}
    lm_cfg = LMConfig(vocab_size=502, hidden=128, n_layers=2, n_heads=2, n_kv_heads=1, head_dim=64, ffn=256)
    pol, ref = B200UnitLM(lm_cfg, device=DEV, max_batch=6, max_seq=40), B200UnitLM(lm_cfg, device=DEV, max_batch=6, max_seq=40, trainable=False)
    pol.load_hf_state_dict(pol_p)
    ref.load_hf_state_dict(ref_p)
    n, T = 3, 40
    ids = torch.randint(2, 502, (2 * n, T), generator=g)
    ids[:, 0] = 1
    ids[n:, :12] = ids[:n, :12]                       # shared prompts
    labels = ids.clone()
    labels[:, :12] = -100
    ids[1, 33:] = 0
    labels[1, 33:] = -100
    loss_o, z_o, grads_o = D.dpo_loss_and_grads(pol_p, ref_p, cfg_o, ids, labels, beta=0.1)
    tr = B200DPOTrainer(pol, ref, beta=0.1, lr=0.0)    # lr 0: gradients stay inspectable after the step
    out = tr.step(ids, labels)
    assert abs(float(out["loss"]) - float(loss_o)) < 2e-3 * abs(float(loss_o)) + 1e-4
    assert float((out["logits_z"].cpu() - z_o).abs().max()) < 5e-3
    sd_g = pol.state_dict_hf(grads=True)
    errs = {k: rel_err(sd_g[k].cpu(), grads_o[k]) for k in pol_p if not k.endswith("k_proj.bias")}
    bad = {k: v for k, v in errs.items() if v > 4e-2}
    assert not bad, bad


def test_dpo_training_moves_the_margin():
    from slamkit_b200.dpo import B200DPOTrainer
    from slamkit_b200.lm import B200UnitLM, LMConfig
    lm_cfg = LMConfig(vocab_size=502, hidden=128, n_layers=2, n_heads=2, n_kv_heads=1, head_dim=64, ffn=256)
    pol = B200UnitLM(lm_cfg, device=DEV, max_batch=8, max_seq=32, seed=0)
    ref = B200UnitLM(lm_cfg, device=DEV, max_batch=8, max_seq=32, seed=0, trainable=False)
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(2, 502, (8, 32), generator=g)
    ids[:, 0] = 1
    ids[4:, :10] = ids[:4, :10]
    labels = ids.clone()
    labels[:, :10] = -100
    tr = B200DPOTrainer(pol, ref, beta=0.1, lr=1e-3)
    first = tr.step(ids, labels)
    assert abs(float(first["loss"]) - 0.6931) < 1e-3          # policy == reference -> z = 0 -> loss = ln 2
    for _ in range(10):
        last = tr.step(ids, labels)
    assert float(last["loss"]) < 0.5 and float(last["logits_z"].mean()) > 0.3
