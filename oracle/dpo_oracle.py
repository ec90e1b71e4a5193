"""CPU oracle for the DPO step.  TEST INFRASTRUCTURE ONLY (same rules as lm_oracle.py).

Restates `trl.DPOTrainer` with `loss_type="sigmoid"`, `beta`, concatenated chosen||rejected forward of the policy and of
a frozen reference model, completion-only log-probabilities, as driven by the reference's `SLAMDPOTrainer`
(slamkit/trainer/slam_dpo_trainer.py:4-64) and cli/preference_alignment_train.py:36-65.  `trl` is not installed in this
image and the reference ships no DPO goldens, so this restatement follows trl's published algorithm and is
**parity-unpinned** against the reference; the GPU path is checked against autograd of this loss."""
from __future__ import annotations

from typing import Dict

import torch

from . import lm_oracle as O


def sequence_logps(p: Dict[str, torch.Tensor], cfg: O.OracleLMConfig, ids: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    logits = O.forward_logits(p, cfg, ids).float()
    lp = torch.log_softmax(logits[:, :-1], dim=-1)
    tgt = labels[:, 1:]
    mask = tgt != -100
    tok = lp.gather(-1, tgt.clamp(min=0).unsqueeze(-1)).squeeze(-1)
    return (tok * mask).sum(-1)


def dpo_loss_and_grads(policy: Dict[str, torch.Tensor], reference: Dict[str, torch.Tensor], cfg: O.OracleLMConfig,
                       ids: torch.Tensor, labels: torch.Tensor, beta: float = 0.1):
    n = ids.shape[0] // 2
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in policy.items()}
    with torch.no_grad():
        ref_lp = sequence_logps(reference, cfg, ids, labels)
    pol_lp = sequence_logps(leaves, cfg, ids, labels)
    z = beta * ((pol_lp[:n] - pol_lp[n:]) - (ref_lp[:n] - ref_lp[n:]))
    loss = -torch.nn.functional.logsigmoid(z).mean()
    loss.backward()
    return loss.detach(), z.detach(), {k: v.grad for k, v in leaves.items()}
