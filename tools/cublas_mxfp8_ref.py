"""Library ceiling for the block-scaled GEMM: cuBLASLt MXFP8 (e4m3 x e4m3, UE8M0 1x32 scales) through
torch._scaled_mm, timed the same way as tools/gemm_time.py.  Context for K3's roofline only -- the
product never calls it."""
import os
import torch

shapes = [tuple(int(v) for v in s.split("x")) for s in os.environ.get("GT_SHAPES", "8192x8192x8192").split(",")]
dev = "cuda"
for (M, N, K) in shapes:
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16).to(torch.float8_e4m3fn)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16).to(torch.float8_e4m3fn)
    rup = lambda x, m: (x + m - 1) // m * m
    sa = torch.full((rup(M, 128) * rup(K // 32, 4),), 127, dtype=torch.uint8, device=dev).view(torch.float8_e8m0fnu)
    sb = torch.full((rup(N, 128) * rup(K // 32, 4),), 127, dtype=torch.uint8, device=dev).view(torch.float8_e8m0fnu)
    try:
        f = lambda: torch._scaled_mm(a, w.t(), scale_a=sa, scale_b=sb, out_dtype=torch.bfloat16)
        for _ in range(3):
            y = f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            y = f()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        ref = a[:64].float() @ w.float().t()
        err = (y[:64].float() - ref).abs().max().item() / ref.abs().max().item()
        print(f"cuBLASLt mxfp8 {M}x{N}x{K}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.0f} TFLOP/s relerr {err:.2e}", flush=True)
    except Exception as e:  # noqa
        print(f"cuBLASLt mxfp8 {M}x{N}x{K}: unavailable ({type(e).__name__}: {str(e)[:200]})", flush=True)
    # plain bf16 for the same shape
    ah, wh = a.to(torch.bfloat16), w.to(torch.bfloat16)
    for _ in range(3):
        torch.nn.functional.linear(ah, wh)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        torch.nn.functional.linear(ah, wh)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"cuBLAS bf16 {M}x{N}x{K}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
