"""Shared helpers for the parity tests."""
import numpy as np
import torch


def u16_to_bf16(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(a.astype(np.uint16)).view(torch.bfloat16)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b|| in fp64."""
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_abs(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).abs().max())
