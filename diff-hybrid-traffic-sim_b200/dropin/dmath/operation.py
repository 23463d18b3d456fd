"""Soft `if`: sigma(clamp(value * constant)).  Interface of the reference's dmath/operation.py:3-30."""
import torch


def sigmoid(value, constant, min=-16.0, max=16.0):
    v = value if isinstance(value, torch.Tensor) else torch.tensor(value)
    c = constant if isinstance(constant, torch.Tensor) else torch.tensor(constant)
    return torch.sigmoid(torch.clamp(v * c, min, max))
