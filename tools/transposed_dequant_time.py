"""`MXTensor.t().to_dtype()` (blocked axis physically innermost, logically second-to-last) and fp32-input `to_mx` on 16384 x 16384:
microseconds and GB/s of algorithmic bytes, CUDA-graph replay."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchmx_b200  # noqa: F401
from torchmx_b200 import dtypes
from torchmx_b200.mx_tensor import MXTensor


def timed(fn, n=4):
    fn(); torch.cuda.synchronize()
    g, st = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(n):
                fn()
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n * 1e3)
    return min(ts[1:])


R = C = 16384
x = torch.randn(R, C, device="cuda", dtype=torch.bfloat16)
out = {}
for name in ("float8_e4m3", "float6_e3m2", "float4_e2m1", "int8"):
    m = MXTensor.to_mx(x, dtypes.STR_TO_ELEM_DTYPE[name], 32)
    per = 0.5 if name == "float4_e2m1" else 1.0
    us_t = timed(lambda: m.t().to_dtype(torch.bfloat16))
    us_f = timed(lambda: m.to_dtype(torch.bfloat16))
    out[name] = {"t().to_dtype(bf16)_us": round(us_t, 1), "GB/s": round(R * C * (2 + per + 1 / 32) / us_t / 1e3, 1), "flat_to_dtype_us": round(us_f, 1)}
xf = x.float()
del x
for name in ("float8_e4m3", "float4_e2m1"):
    per = 0.5 if name == "float4_e2m1" else 1.0
    us = timed(lambda: MXTensor.to_mx(xf, dtypes.STR_TO_ELEM_DTYPE[name], 32))
    out[f"to_mx_fp32_input_{name}"] = {"us": round(us, 1), "GB/s": round(R * C * (4 + per + 1 / 32) / us / 1e3, 1)}
print(json.dumps(out, indent=1))
