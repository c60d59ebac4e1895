"""SQNR helper the reference's tests import (tests/test_mx_tensor.py:13)."""
import torch


def compute_error(x, y):
    return 20 * torch.log10(torch.linalg.norm(x) / torch.linalg.norm(x - y))
