"""One eager decode step (batch 32) of the HF Llama-3-8B shape (2 layers) quantized with quantize_llm_(fuse_rmsnorm=True), for an
ncu launch list: which kernels a decoder layer launches between the MX linears."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_bench as lb
from transformers.cache_utils import StaticCache
with torch.no_grad():
    model, cfg, info = lb.build("8b", 2, "float6_e3m2", "float8_e4m3", llm_api=True, fuse_norm=True, mx_attention=os.environ.get("DB_MX_ATTENTION", "0") == "1")
    B, ctx = 32, 128
    cache = StaticCache(config=cfg, max_cache_len=256 if os.environ.get("DB_MX_ATTENTION", "0") == "1" else ctx + 16)
    model(input_ids=torch.randint(0, cfg.vocab_size, (B, ctx), device="cuda"), past_key_values=cache, use_cache=True)
    tok = torch.randint(0, cfg.vocab_size, (B, 1), device="cuda")
    for i in range(3):
        model(input_ids=tok, past_key_values=cache, use_cache=True)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    model(input_ids=tok, past_key_values=cache, use_cache=True)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done")
