"""K1b (silu * up -> MX codes) at a size well past L2: us per launch and GB/s of algorithmic bytes (2 + 2 + 1 + 1/32 per element)."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchmx_b200  # noqa
from torchmx_b200 import dtypes, mlp_ops
rows, cols = int(os.environ.get("K1B_ROWS", "16384")), int(os.environ.get("K1B_COLS", "14336"))
out = {}
for name in os.environ.get("K1B_ELEMS", "float8_e4m3,float6_e3m2,float4_e2m1,int8").split(","):
    el = getattr(dtypes, name)
    g = torch.randn(rows, cols, device="cuda", dtype=torch.bfloat16) * 2
    u = torch.randn(rows, cols, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        mlp_ops.silu_mul_to_mx(g, u, el)
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); mlp_ops.silu_mul_to_mx(g, u, el); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    by = rows * cols * (4 + (0.5 if name == "float4_e2m1" else 1) + 1 / 32)
    out[name] = {"us": round(us, 1), "GBps": round(by / us / 1e3, 1)}
print(json.dumps({"rows": rows, "cols": cols, "k1b": out}))
