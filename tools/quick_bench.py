"""Developer micro-benchmark: device-resident quantize / dequantize at 16384^2 through the C ABI."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchmx_b200  # noqa
from torchmx_b200 import dtypes

N = int(os.environ.get("QB_N", 16384))
ITERS = int(os.environ.get("QB_ITERS", 20))
elems = os.environ.get("QB_ELEMS", "float8_e4m3,float6_e3m2,float6_e2m3,float4_e2m1,int8").split(",")
dev = "cuda:0"
xs = [torch.randn(N, N, dtype=torch.bfloat16, device=dev) for _ in range(3)]


def timeit(fn, iters=ITERS):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i in range(iters):
        ev[i][0].record()
        fn(i)
        ev[i][1].record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e-3, ts[0] * 1e-3


print(f"ept={os.environ.get('MXQ_QUANT_EPT', 'default')} N={N}")
for e in elems:
    per = 0.5 if e == "float4_e2m1" else 1.0
    qbytes = N * N * (2 + per + 1 / 32)
    med, best = timeit(lambda i: torch.ops.torchmx.quantize_mx(xs[i % 3], e, 32))
    s, c = torch.ops.torchmx.quantize_mx(xs[0], e, 32)
    dmed, dbest = timeit(lambda i: torch.ops.torchmx.dequantize_mx(c, s, e, 32, torch.bfloat16, 1))
    fbytes = N * N * (4 + per + 1 / 32)
    fmed, fbest = timeit(lambda i: torch.ops.torchmx.dequantize_mx(c, s, e, 32, torch.float32, 1))
    print(f"{e:12s} quant {qbytes/med/1e9:7.0f} GB/s (best {qbytes/best/1e9:6.0f}, {med*1e6:6.1f} us) | dequant->bf16 {qbytes/dmed/1e9:7.0f} (best {qbytes/dbest/1e9:6.0f}, {dmed*1e6:6.1f} us)"
          f" | ->f32 {fbytes/fmed/1e9:7.0f} (best {fbytes/fbest/1e9:6.0f}, {fmed*1e6:6.1f} us)", flush=True)
a = xs[0]; b = torch.empty_like(a)
med, best = timeit(lambda i: b.copy_(a))
print(f"torch copy_ bf16: {2*a.numel()*2/med/1e9:.0f} GB/s (best {2*a.numel()*2/best/1e9:.0f})")
