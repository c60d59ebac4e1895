"""One launch per GEMM variant (for an ncu metrics pass): GT_CFGS list as in gemm_time.py, plus 'cublas'."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchmx_b200  # noqa
from torchmx_b200 import dtypes
from torchmx_b200.mx_tensor import MXTensor
for shape in os.environ.get("GT_SHAPES", "8192x8192x8192").split(","):
  M, N, K = (int(v) for v in shape.split("x"))
  a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
  b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
  A = MXTensor.to_mx(a, dtypes.float8_e4m3, 32)
  B = MXTensor.to_mx(b, dtypes.float6_e3m2, 32)
  a8, b8 = a.to(torch.float8_e4m3fn), b.to(torch.float8_e4m3fn)
  rup = lambda x, m: (x + m - 1) // m * m
  sa = torch.full((rup(M, 128) * rup(K // 32, 4),), 127, dtype=torch.uint8, device="cuda").view(torch.float8_e8m0fnu)
  sb = torch.full((rup(N, 128) * rup(K // 32, 4),), 127, dtype=torch.uint8, device="cuda").view(torch.float8_e8m0fnu)
  for rep in range(int(os.environ.get("GT_REPS", "2"))):
      for c in os.environ.get("GT_CFGS", "0").split(","):
          if c == "cublas":
              y = torch._scaled_mm(a8, b8.t(), scale_a=sa, scale_b=sb, out_dtype=torch.bfloat16)
          else:
              cfg, _, gm = c.partition(":")
              os.environ["MXQ_GEMM_CFG"] = cfg
              os.environ["MXQ_GEMM_GM"] = gm or "0"
              y = torch.nn.functional.linear(A, B)
          torch.cuda.synchronize()
print("done")
