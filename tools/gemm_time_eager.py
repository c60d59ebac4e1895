"""Single-launch timing (events around one launch, device idle before): GT_SHAPES list."""
import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchmx_b200  # noqa
from torchmx_b200 import dtypes
from torchmx_b200.mx_tensor import MXTensor
for shape in os.environ.get("GT_SHAPES", "32x128256x4096").split(","):
    M, N, K = (int(v) for v in shape.split("x"))
    A = MXTensor.to_mx(torch.randn(M, K, device="cuda", dtype=torch.bfloat16), dtypes.float8_e4m3, 32)
    B = MXTensor.to_mx(torch.randn(N, K, device="cuda", dtype=torch.bfloat16), dtypes.float6_e3m2, 32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for mode in ("idle", "back2back", "flushed"):
        ts = []
        for i in range(12):
            if mode == "flushed":
                flush.zero_()
            if mode != "back2back":
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y = torch.nn.functional.linear(A, B)
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1) * 1e3)
        from torchmx_b200 import mx_gemm
        print(mx_gemm.stats, end=" ")
        print(f"{shape} {mode}: min {min(ts):.1f} us median {statistics.median(ts):.1f} us", flush=True)
