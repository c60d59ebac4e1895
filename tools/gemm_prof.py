import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchmx_b200  # noqa
from torchmx_b200 import dtypes
from torchmx_b200.mx_tensor import MXTensor
M = N = K = int(os.environ.get("GP_N", 8192))
a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
A = MXTensor.to_mx(a, dtypes.float8_e4m3, 32)
B = MXTensor.to_mx(b, dtypes.float6_e3m2, 32)
for _ in range(int(os.environ.get("GP_ITERS", 3))):
    y = torch.nn.functional.linear(A, B)
torch.cuda.synchronize()
print("ok", y.float().abs().mean().item())
