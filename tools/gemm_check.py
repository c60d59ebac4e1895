"""Developer check of the tensor-core MX GEMM against dequantize-then-matmul in fp64."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchmx_b200  # noqa
from torchmx_b200 import dtypes, mx_gemm
from torchmx_b200.mx_tensor import MXTensor

dev = "cuda:0"
torch.manual_seed(0)


def check(M, N, K, ea, eb, bias=False, batch=0, scale_spread=0):
    shape_a = (batch, M, K) if batch else (M, K)
    shape_b = (batch, N, K) if batch else (N, K)
    a = torch.randn(*shape_a, device=dev, dtype=torch.bfloat16)
    b = torch.randn(*shape_b, device=dev, dtype=torch.bfloat16)
    if scale_spread:
        a *= torch.exp2(torch.randint(-scale_spread, scale_spread, (*shape_a[:-1], K // 32), device=dev).float()).repeat_interleave(32, -1).to(torch.bfloat16)
        b *= torch.exp2(torch.randint(-scale_spread, scale_spread, (*shape_b[:-1], K // 32), device=dev).float()).repeat_interleave(32, -1).to(torch.bfloat16)
    A = MXTensor.to_mx(a, getattr(dtypes, ea), 32)
    B = MXTensor.to_mx(b, getattr(dtypes, eb), 32)
    bias_t = torch.randn(N, device=dev, dtype=torch.bfloat16) if bias else None
    mx_gemm.stats["tensor_core"] = mx_gemm.stats["fallback"] = 0
    if batch:
        out = torch.bmm(A, B.transpose(1, 2))
    else:
        out = torch.nn.functional.linear(A, B, bias_t)
    torch.cuda.synchronize()
    used_tc = mx_gemm.stats["tensor_core"] > 0
    ad, bd = A.to_dtype(torch.float32).double(), B.to_dtype(torch.float32).double()
    ref = ad @ bd.transpose(-1, -2)
    S = ad.abs() @ bd.abs().transpose(-1, -2)
    if bias:
        ref = ref + bias_t.double()
    err = (out.double() - ref).abs()
    tol = 2.0 ** -8 * ref.abs() + 2.0 ** -18 * S + 1e-30
    bad = (err > tol).sum().item()
    rel = (err / (S + 1e-30)).max().item()
    print(f"M={M} N={N} K={K} {ea}x{eb} bias={bias} batch={batch} spread={scale_spread}: tc={used_tc} bad={bad}/{err.numel()} max err/S={rel:.3e} "
          f"nan={torch.isnan(out).sum().item()}", flush=True)
    return bad == 0 and used_tc


ok = True
ok &= check(128, 128, 128, "float8_e4m3", "float8_e4m3")
ok &= check(128, 256, 256, "float8_e4m3", "float8_e4m3")
ok &= check(256, 512, 512, "float8_e4m3", "float6_e3m2", scale_spread=8)
ok &= check(100, 200, 384, "float8_e4m3", "float4_e2m1", bias=True)
ok &= check(1, 4096, 4096, "float8_e4m3", "float6_e3m2")
ok &= check(33, 130, 1024, "float6_e2m3", "float6_e3m2", bias=True, scale_spread=20)
ok &= check(2048, 2048, 128, "float8_e4m3", "float6_e3m2", batch=4)
ok &= check(77, 300, 256, "float4_e2m1", "float4_e2m1", batch=3)
ok &= check(1024, 4096, 4096, "float8_e4m3", "float6_e3m2", scale_spread=4)
# decode-sized activations: the skinny weight-streaming kernel (K3c), every token-tile width, auto and forced K splits
for forced in ("", "1", "2", "8"):
    mx_gemm.overrides["split_k"] = int(forced or 0)
    print(f"-- skinny, split_k={forced or 'auto'}")
    ok &= check(32, 4096, 4096, "float8_e4m3", "float6_e3m2", bias=True, scale_spread=6)
    ok &= check(17, 1000, 1024, "float8_e4m3", "float4_e2m1")
    ok &= check(64, 1024, 4096, "float8_e4m3", "float6_e3m2", bias=True)
    ok &= check(100, 384, 2048, "float6_e2m3", "float6_e3m2", scale_spread=12)
    ok &= check(128, 14336, 4096, "float8_e4m3", "float6_e3m2")
mx_gemm.overrides["split_k"] = 0
ok &= check(32, 128256, 4096, "float8_e4m3", "float6_e3m2")
ok &= check(5, 4096, 14336, "float8_e4m3", "float6_e3m2", bias=True)
print("ALL OK" if ok else "FAILURES")

if ok and os.environ.get("GEMM_BENCH", "1") == "1":
    for (M, N, K) in [(8192, 8192, 8192), (2048, 14336, 4096), (32, 4096, 4096)]:
        a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        b = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
        A = MXTensor.to_mx(a, dtypes.float8_e4m3, 32)
        B = MXTensor.to_mx(b, dtypes.float6_e3m2, 32)
        for _ in range(3):
            torch.nn.functional.linear(A, B)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            torch.nn.functional.linear(A, B)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"linear {M}x{N}x{K}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
        ah, bh = A.to_dtype(torch.bfloat16), B.to_dtype(torch.bfloat16)
        for _ in range(3):
            torch.nn.functional.linear(ah, bh)
        e0.record()
        for _ in range(n):
            torch.nn.functional.linear(ah, bh)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"  cuBLAS bf16 on dequantized operands: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
