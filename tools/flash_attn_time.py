"""K4b (one kernel) vs the chain it replaces (K3 bmm -> K4a -> K3 bmm + transpose) on the attention of a Llama-3-8B layer at
2048 tokens: 32 query heads, 8 key / value heads, head_dim 128, e4m3 Q / K / V / P.  CUDA-graph replay, microseconds."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import torchmx_b200  # noqa: F401
from torchmx_b200 import attention_ops, dtypes
from torchmx_b200.layers.mx_llama_attention import _repeat_heads
from torchmx_b200.mx_tensor import MXTensor


def timed(fn, n=10):
    fn()
    torch.cuda.synchronize()
    g, st = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(n):
                fn()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n * 1e3)
    return round(min(ts[1:]), 1)


once = len(sys.argv) > 1 and sys.argv[1] == "--once"  # one launch of each variant (for ncu)
out = {}
for (B, H, HK, S) in ((1, 32, 8, 2048), (1, 32, 8, 4096), (8, 32, 8, 512)):
    q = torch.randn(B, H, S, 128, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(B, HK, S, 128, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(B, HK, S, 128, device="cuda", dtype=torch.bfloat16)
    E = dtypes.float8_e4m3
    Q, K, VT = MXTensor.to_mx(q, E, 32), MXTensor.to_mx(k, E, 32), MXTensor.to_mx(v.transpose(2, 3).contiguous(), E, 32)
    sc = 128 ** -0.5
    addm = torch.zeros(1, 1, S, S, device="cuda", dtype=torch.bfloat16).masked_fill_(torch.ones(S, S, dtype=torch.bool, device="cuda").triu_(1), torch.finfo(torch.bfloat16).min)

    def chain(mask, causal):
        kk, vv = _repeat_heads(K, H // HK), _repeat_heads(VT, H // HK).transpose(2, 3)
        s = torch.matmul(Q, kk.transpose(2, 3))
        p = attention_ops.softmax_to_mx(s, sc, mask, causal, E, 32)
        return torch.matmul(p, vv).transpose(1, 2).contiguous()

    if once:
        attention_ops.flash_attention(Q, K, VT, sc, None, True, E, 32)
        attention_ops.flash_attention(Q, K, VT, sc, addm, False, E, 32)
        torch.cuda.synchronize()
        break
    flops = 4.0 * B * H * S * S * 128  # both contractions, every key
    r = {}
    r["flash_causal_us"] = timed(lambda: attention_ops.flash_attention(Q, K, VT, sc, None, True, E, 32))
    r["flash_additive_mask_us"] = timed(lambda: attention_ops.flash_attention(Q, K, VT, sc, addm, False, E, 32))
    r["chain_causal_us"] = timed(lambda: chain(None, True), n=4)
    r["chain_additive_mask_us"] = timed(lambda: chain(addm, False), n=4)
    r["sdpa_bf16_causal_us"] = timed(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=True))
    r["algorithmic_bytes"] = int((B * H * S * 128 + 2 * B * HK * S * 128) * (1 + 1 / 32) + B * H * S * 128 * 2)
    r["chain_extra_hbm_bytes"] = int(B * H * S * S * (2 + 2 + 1 + 1 / 32 + 1 + 1 / 32))  # scores written + read, P written + read
    r["flash_causal_TFLOPs_of_visible_work"] = round(flops / 2 * 3 / 2 / r["flash_causal_us"] / 1e6, 1)  # 3 Q K^T passes + 1 P V over the causal half
    out[f"b{B}_h{H}_kv{HK}_s{S}"] = r
    del q, k, v, Q, K, VT, addm
if not once:
    print(json.dumps(out, indent=1))
