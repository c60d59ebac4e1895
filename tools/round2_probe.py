"""Timings of the round-2 kernels through the public ops, CUDA-graph replay (host dispatch not in the number):
kind::mxf4 vs kind::mxf8f6f4 on fp4 x fp4, the fused dequantize GEMM (K3d) vs the K2 + cuBLAS recipe it replaces."""
import json
import statistics
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import torchmx_b200  # noqa: F401
from torchmx_b200 import dtypes, mx_gemm
from torchmx_b200.mx_tensor import MXTensor

dev = torch.device("cuda:0")


def timed(fn, n=10, rounds=5):
    fn()
    torch.cuda.synchronize()
    g, st = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(n):
                fn()
    ts = []
    for r in range(rounds + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        if r:
            ts.append(e0.elapsed_time(e1) / n * 1e3)
    return round(min(ts), 1), round(statistics.median(ts), 1)


out = {}
gen = torch.Generator(device=dev).manual_seed(0)
E = dtypes.STR_TO_ELEM_DTYPE
for (M, N, K) in ((8192, 8192, 8192), (4096, 14336, 4096), (2048, 4096, 4096)):
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16, generator=gen)
    b = torch.randn(N, K, device=dev, dtype=torch.bfloat16, generator=gen)
    flops = 2.0 * M * N * K
    for ea, eb in (("float8_e4m3", "float6_e3m2"), ("float4_e2m1", "float4_e2m1")):
        A, B = MXTensor.to_mx(a, E[ea], 32), MXTensor.to_mx(b, E[eb], 32)
        for no_mxf4 in ((False, True) if ea == "float4_e2m1" else (False,)):
            mx_gemm.overrides["no_mxf4"] = no_mxf4
            best, med = timed(lambda: torch.nn.functional.linear(A, B))
            out[f"linear_{M}x{N}x{K}_{ea}x{eb}" + ("_mxf8f6f4" if no_mxf4 else "")] = {"us": med, "us_best": best, "TFLOP/s": round(flops / med / 1e6, 1)}
        mx_gemm.overrides["no_mxf4"] = False
    # int8: K3d vs dequantize + cuBLAS
    A, B = MXTensor.to_mx(a, E["int8"], 32), MXTensor.to_mx(b, E["int8"], 32)
    best, med = timed(lambda: torch.nn.functional.linear(A, B), n=4)
    out[f"linear_{M}x{N}x{K}_int8_k3d"] = {"us": med, "us_best": best, "TFLOP/s": round(flops / med / 1e6, 1)}
    prev = mx_gemm.set_dequant_gemm(False)
    best, med = timed(lambda: torch.nn.functional.linear(A, B), n=4)
    mx_gemm.set_dequant_gemm(prev)
    out[f"linear_{M}x{N}x{K}_int8_k2_plus_cublas"] = {"us": med, "us_best": best, "TFLOP/s": round(flops / med / 1e6, 1)}
    del a, b, A, B
# README-sized matmul (B blocked along N) and a decode-sized int8 linear
x = MXTensor.to_mx(torch.randn(128, 128, device=dev, dtype=torch.bfloat16), E["float8_e4m3"], 32)
y = MXTensor.to_mx(torch.randn(128, 128, device=dev, dtype=torch.bfloat16), E["float6_e3m2"], 32)
out["readme_matmul_128"] = dict(zip(("us_best", "us"), timed(lambda: torch.matmul(x, y))))
X = MXTensor.to_mx(torch.randn(32, 4096, device=dev, dtype=torch.bfloat16), E["int8"], 32)
W = MXTensor.to_mx(torch.randn(14336, 4096, device=dev, dtype=torch.bfloat16), E["int8"], 32)
best, med = timed(lambda: torch.nn.functional.linear(X, W))
out["linear_32x14336x4096_int8_k3d"] = {"us": med, "us_best": best, "GB/s": round(14336 * 4096 * (1 + 1 / 32) / med / 1e3, 1)}
out["stats"] = dict(mx_gemm.stats)
print(json.dumps(out, indent=1))
