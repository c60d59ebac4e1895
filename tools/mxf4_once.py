"""one kind::mxf4 launch (fp4 x fp4, 8192^3) for an ncu capture"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchmx_b200  # noqa: F401
from torchmx_b200 import dtypes
from torchmx_b200.mx_tensor import MXTensor
a = MXTensor.to_mx(torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16), dtypes.float4_e2m1, 32)
w = MXTensor.to_mx(torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16), dtypes.float4_e2m1, 32)
for _ in range(3):
    torch.nn.functional.linear(a, w)
torch.cuda.synchronize()
