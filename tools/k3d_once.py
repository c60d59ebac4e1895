"""one K3d launch (int8 x int8, 2048 x 4096 x 4096) for an ncu capture"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import torchmx_b200  # noqa: F401
from torchmx_b200 import dtypes
from torchmx_b200.mx_tensor import MXTensor

a = MXTensor.to_mx(torch.randn(2048, 4096, device="cuda", dtype=torch.bfloat16), dtypes.int8, 32)
w = MXTensor.to_mx(torch.randn(4096, 4096, device="cuda", dtype=torch.bfloat16), dtypes.int8, 32)
for _ in range(2):
    torch.nn.functional.linear(a, w)
torch.cuda.synchronize()
