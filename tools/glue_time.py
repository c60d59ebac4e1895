"""K5a-d / K1b at Llama-3-8B prefill (2048 tokens) and decode (batch 32, 256-slot cache) shapes: us per launch (CUDA-graph replay of 20
launches) and GB/s of algorithmic bytes."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchmx_b200  # noqa
from torchmx_b200 import dtypes, glue_ops, mlp_ops
e8 = dtypes.float8_e4m3
def timed(f):
    for _ in range(3): f()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): f()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 20 * 1e3)
    return min(ts)
out = {}
bf = lambda *s: torch.randn(*s, device="cuda", dtype=torch.bfloat16)
for name, rows in (("prefill_2048", 2048), ("decode_32", 32)):
    b, t = (1, rows) if rows > 32 else (rows, 1)
    x, r, w = bf(b, t, 4096), bf(b, t, 4096), bf(4096)
    us = timed(lambda: glue_ops.rmsnorm(x, w, 1e-5, residual=r, to_mx=e8, want_y=False))
    by = rows * 4096 * (2 + 2 + 2 + 1 + 1 / 32)
    out[f"K5a_rmsnorm_residual_to_mx_{name}"] = {"us": round(us, 2), "GB/s": round(by / us / 1e3)}
    qkv = bf(b, t, 6144)
    q, k, v = (z.view(b, t, -1, 128).transpose(1, 2) for z in qkv.split([4096, 1024, 1024], -1))
    cs = bf(1, t, 128)
    us = timed(lambda: glue_ops.rope(q, k, cs, cs))
    by = rows * 5120 * 4 + t * 128 * 4
    out[f"K5b_rope_{name}"] = {"us": round(us, 2), "GB/s": round(by / us / 1e3)}
    kc, vc = bf(b, 8, max(t, 256), 128), bf(b, 8, max(t, 256), 128)
    us = timed(lambda: glue_ops.rope(q, k, cs, cs, k_out=kc[:, :, :t], v=v, v_out=vc[:, :, :t]))
    out[f"K5b_rope_with_cache_write_{name}"] = {"us": round(us, 2), "GB/s": round((by + rows * 1024 * 4) / us / 1e3)}
    a = bf(b, 32, t, 128)
    us = timed(lambda: glue_ops.quantize_heads(a, e8))
    out[f"K5c_quantize_heads_{name}"] = {"us": round(us, 2), "GB/s": round(rows * 4096 * (3 + 1 / 32) / us / 1e3)}
    kv = 2048 if rows > 32 else 256
    vv = bf(b, 8, kv, 128)
    us = timed(lambda: glue_ops.quantize_transposed(vv, e8))
    out[f"K5d_quantize_transposed_{name}"] = {"us": round(us, 2), "GB/s": round(b * 8 * kv * 128 * (3 + 1 / 32) / us / 1e3)}
    gu = bf(b, t, 2 * 14336)
    g_, u_ = gu.split([14336, 14336], -1)
    us = timed(lambda: mlp_ops.silu_mul_to_mx(g_, u_, e8, 32))
    out[f"K1b_silu_mul_to_mx_{name}"] = {"us": round(us, 2), "GB/s": round(rows * 14336 * (5 + 1 / 32) / us / 1e3)}
print(json.dumps(out, indent=1))
