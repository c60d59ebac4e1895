"""Kernel-level timing of the MX GEMM through the public op (F.linear on MXTensors), replayed from a CUDA
graph so host dispatch is not in the number.  GT_SHAPES=MxNxK,..  GT_CFGS=comma list of MXQ_GEMM_CFG[:GM] values
(0 = default dispatch), interleaved over GT_ROUNDS rounds; prints min / median per config."""
import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchmx_b200  # noqa
from torchmx_b200 import dtypes
from torchmx_b200.mx_tensor import MXTensor
shapes = [tuple(int(v) for v in s.split("x")) for s in os.environ.get("GT_SHAPES", "8192x8192x8192").split(",")]
cfgs = os.environ.get("GT_CFGS", "0").split(",")
rounds = int(os.environ.get("GT_ROUNDS", "3"))
n = int(os.environ.get("GT_ITERS", "10"))
wdt = getattr(dtypes, os.environ.get("GT_W", "float6_e3m2"))
for (M, N, K) in shapes:
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    A = MXTensor.to_mx(a, dtypes.float8_e4m3, 32)
    B = MXTensor.to_mx(b, wdt, 32)
    ref = A.to_dtype(torch.float32)[:64] @ B.to_dtype(torch.float32).t()
    graphs = {}
    a8, b8 = a.to(torch.float8_e4m3fn), b.to(torch.float8_e4m3fn)
    rup = lambda x, m: (x + m - 1) // m * m
    sa = torch.full((rup(M, 128) * rup(K // 32, 4),), 127, dtype=torch.uint8, device="cuda").view(torch.float8_e8m0fnu)
    sb = torch.full((rup(N, 128) * rup(K // 32, 4),), 127, dtype=torch.uint8, device="cuda").view(torch.float8_e8m0fnu)
    for c in cfgs:
        if c == "cublas":  # library ceiling (cuBLASLt MXFP8), context only
            f = lambda: torch._scaled_mm(a8, b8.t(), scale_a=sa, scale_b=sb, out_dtype=torch.bfloat16)
            f(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph(); st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                with torch.cuda.graph(g, stream=st):
                    for _ in range(n):
                        y = f()
            graphs[c] = (g, float("nan"))
            continue
        cfg, _, gm = c.partition(":")
        os.environ["MXQ_GEMM_CFG"] = cfg
        os.environ["MXQ_GEMM_GM"] = gm or "0"
        y = torch.nn.functional.linear(A, B)
        torch.cuda.synchronize()
        err = (y[:64].float() - ref).abs().max().item() / ref.abs().max().item()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for _ in range(n):
                    y = torch.nn.functional.linear(A, B)
        graphs[c] = (g, err)
    times = {c: [] for c in cfgs}
    for r in range(rounds + 1):
        for c in cfgs:
            g, _ = graphs[c]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            if r > 0:
                times[c].append(e0.elapsed_time(e1) / n * 1e3)
    for c in cfgs:
        t = times[c]
        wbytes = N * K * (1 + 1 / 32) + M * K * (1 + 1 / 32) + M * N * 2
        print(f"{M}x{N}x{K} cfg={c}: min {min(t):.1f} us ({2*M*N*K/min(t)/1e6:.0f} TFLOP/s, {wbytes/min(t)/1e3:.0f} GB/s operand+output bytes) median {statistics.median(t):.1f} us relerr {graphs[c][1]:.2e}", flush=True)
