"""One K1b launch on a 2048-token Llama-3-8B MLP activation (column slices of a stacked gate+up output), for ncu."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchmx_b200  # noqa
from torchmx_b200 import dtypes, mlp_ops
gu = torch.randn(2048, 2 * 14336, device="cuda", dtype=torch.bfloat16)
g, u = gu.split([14336, 14336], dim=-1)
for _ in range(3):
    mlp_ops.silu_mul_to_mx(g, u, dtypes.float8_e4m3, 32)
torch.cuda.synchronize()
print("ok")
