"""One eager decode step of the quantized HF Llama (8B shape, few layers) for an ncu launch list: which kernels run."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_bench as lb
from transformers.cache_utils import StaticCache
with torch.no_grad():
    model, cfg, info = lb.build("8b", 2, "float6_e3m2", "float8_e4m3")
    B, ctx = 32, 128
    cache = StaticCache(config=cfg, max_cache_len=ctx + 16)
    model(input_ids=torch.randint(0, cfg.vocab_size, (B, ctx), device="cuda"), past_key_values=cache, use_cache=True)
    tok = torch.randint(0, cfg.vocab_size, (B, 1), device="cuda")
    for i in range(3):
        torch.cuda.synchronize()
        if i == 2:
            torch.cuda.nvtx.range_push("decode_step")
        model(input_ids=tok, past_key_values=cache, use_cache=True)
        if i == 2:
            torch.cuda.nvtx.range_pop()
    torch.cuda.synchronize()
print("done")
