import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchmx_b200  # noqa
from torchmx_b200 import dtypes
from torchmx_b200.mx_tensor import MXTensor
shapes = [tuple(int(v) for v in s.split("x")) for s in os.environ.get("GT_SHAPES", "8192x8192x8192").split(",")]
for (M, N, K) in shapes:
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    A = MXTensor.to_mx(a, dtypes.float8_e4m3, 32)
    B = MXTensor.to_mx(b, dtypes.float6_e3m2, 32)
    for _ in range(3):
        y = torch.nn.functional.linear(A, B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        y = torch.nn.functional.linear(A, B)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    ref = A.to_dtype(torch.float32)[:64] @ B.to_dtype(torch.float32).t()
    err = (y[:64].float() - ref).abs().max().item() / ref.abs().max().item()
    print(f"cfg={os.environ.get('MXQ_GEMM_CFG','-')} narrow={os.environ.get('MXQ_GEMM_NARROW','0')} {M}x{N}x{K}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.0f} TFLOP/s relerr {err:.2e}", flush=True)
