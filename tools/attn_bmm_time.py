"""Attention-shaped 4-D MX matmuls (BASELINE configs[2], second half): Q@K^T and P@V for [1, 32, 2048, 128] heads."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchmx_b200  # noqa
from torchmx_b200 import dtypes, mx_gemm
from torchmx_b200.mx_tensor import MXTensor
def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    g, st = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(n): fn()
    ts = []
    for r in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n * 1e3)
    return min(ts[1:])
B, H, S, D = 1, 32, 2048, 128
q = torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16); k = torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16)
v = torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16); p = torch.softmax(torch.randn(B, H, S, S, device="cuda"), -1).to(torch.bfloat16)
Q, K = MXTensor.to_mx(q, dtypes.float8_e4m3, 32), MXTensor.to_mx(k, dtypes.float8_e4m3, 32)
P = MXTensor.to_mx(p, dtypes.float8_e4m3, 32)
V = MXTensor.to_mx(v.transpose(2, 3).contiguous(), dtypes.float8_e4m3, 32).transpose(2, 3)
s0 = dict(mx_gemm.stats)
t_qk = timed(lambda: torch.matmul(Q, K.transpose(2, 3)))
t_pv = timed(lambda: torch.matmul(P, V))
print("tensor-core calls:", mx_gemm.stats["tensor_core"] - s0["tensor_core"], "fallback:", mx_gemm.stats["fallback"] - s0["fallback"])
print(f"MX Q@K^T: {t_qk:.1f} us ({H*S*S*2/t_qk/1e3:.0f} GB/s of bf16 output); cuBLAS bf16: {timed(lambda: torch.matmul(q, k.transpose(2, 3))):.1f} us")
print(f"MX P@V:   {t_pv:.1f} us ({(H*S*S*(1+1/32))/t_pv/1e3:.0f} GB/s of fp8 P read); cuBLAS bf16: {timed(lambda: torch.matmul(p, v)):.1f} us")
t_qp = timed(lambda: MXTensor.to_mx(p, dtypes.float8_e4m3, 32))
print(f"to_mx(P) [{H}x{S}x{S}]: {t_qp:.1f} us ({H*S*S*(3+1/32)/t_qp/1e3:.0f} GB/s)")

# the chain between the two matmuls (reference mx_llama_attention.py:214-239): unfused aten ops + K1 vs the fused K4a kernel
from torchmx_b200 import attention_ops
scores = torch.matmul(Q, K.transpose(2, 3))
sc = D ** -0.5
def unfused():
    w = scores * sc
    w = w.masked_fill(hidden, float("-inf"))
    w = torch.softmax(w, dim=-1, dtype=torch.float32).to(torch.bfloat16)
    return MXTensor.to_mx(w, dtypes.float8_e4m3, 32)
hidden = torch.ones(S, S, dtype=torch.bool, device="cuda").triu_(1)
t_un = timed(unfused, n=4)
t_fu = timed(lambda: attention_ops.softmax_to_mx(scores, sc, None, True, dtypes.float8_e4m3, 32))
print(f"scale+mask+softmax+to_mx(P): unfused {t_un:.1f} us, fused K4a {t_fu:.1f} us ({H*S*S*(3+1/32)/t_fu/1e3:.0f} GB/s of its 3.03 B/elem)")
addm = torch.zeros(1, 1, S, S, device="cuda", dtype=torch.bfloat16).masked_fill_(hidden, float("-inf"))
t_fm = timed(lambda: attention_ops.softmax_to_mx(scores, sc, addm, False, dtypes.float8_e4m3, 32))
t_fn = timed(lambda: attention_ops.softmax_to_mx(scores, sc, None, False, dtypes.float8_e4m3, 32))
print(f"fused K4a with an explicit additive causal mask: {t_fm:.1f} us; no mask at all (every block live): {t_fn:.1f} us ({H*S*S*(3+1/32)/t_fn/1e3:.0f} GB/s)")
