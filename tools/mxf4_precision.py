"""How exactly do kind::mxf4 and kind::mxf8f6f4 accumulate?  Worst |out - ref| / sum|a_k b_k| over fp4 x fp4 operands whose block
scales spread over 2^+-spread, for both instruction kinds (same operand bytes), K = 256 .. 8192."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import torchmx_b200  # noqa: F401
from torchmx_b200 import dtypes, mx_gemm
from torchmx_b200.mx_tensor import MXTensor

dev = torch.device("cuda:0")
mx_gemm.overrides["wide_tiles"] = True
res = []
for K in (256, 768, 2048, 8192):
    for spread in (0, 4, 10, 20):
        g = torch.Generator(device=dev).manual_seed(K + spread)
        a = torch.randn(512, K, device=dev, dtype=torch.bfloat16, generator=g)
        b = torch.randn(512, K, device=dev, dtype=torch.bfloat16, generator=g)
        if spread:
            for t in (a, b):
                e = torch.randint(-spread, spread, (512, K // 32), device=dev, generator=g).float()
                t *= torch.exp2(e).repeat_interleave(32, -1).to(torch.bfloat16)
        A, B = MXTensor.to_mx(a, dtypes.float4_e2m1, 32), MXTensor.to_mx(b, dtypes.float4_e2m1, 32)
        ad, bd = A.to_dtype(torch.float32).double(), B.to_dtype(torch.float32).double()
        ref, S = ad @ bd.t(), ad.abs() @ bd.abs().t()
        row = {"K": K, "spread": spread}
        for name, flag in (("mxf4", False), ("mxf8f6f4", True)):
            mx_gemm.overrides["no_mxf4"] = flag
            out = torch.nn.functional.linear(A, B).double()
            err = (out - ref).abs()
            excess = (err - 2.0 ** -8 * ref.abs()).clamp(min=0)  # what is left after one bf16 ulp of the result
            row[name] = {"worst_excess_over_S_log2": round(float(torch.log2((excess / S).max() + 1e-300)), 2),
                         "outside_2^-18": int((excess > 2.0 ** -18 * S).sum()), "outside_2^-14": int((excess > 2.0 ** -14 * S).sum())}
        mx_gemm.overrides["no_mxf4"] = False
        res.append(row)
        print(json.dumps(row), flush=True)
