"""Kernel-time breakdown of one MX-Llama prefill (few layers) with the MX attention block: which launches the time goes to."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tools.llama_bench as lb
from torch.profiler import profile, ProfilerActivity
layers = int(os.environ.get("PB_LAYERS", "4"))
mode = os.environ.get("PB_MODE", "mx_attention")
model, cfg, info = lb.build("8b", layers, "float6_e3m2", "float8_e4m3", llm_api=mode != "linear", mx_attention=mode == "mx_attention")
ids = torch.randint(0, cfg.vocab_size, (1, 2048), device="cuda")
from transformers.cache_utils import StaticCache
cache = StaticCache(config=cfg, max_cache_len=2176)
with torch.no_grad():
    for _ in range(2):
        lb.set_len(cache, 0); model(input_ids=ids, past_key_values=cache, use_cache=True)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        lb.set_len(cache, 0); model(input_ids=ids, past_key_values=cache, use_cache=True)
        torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time {tot/1e3:.2f} ms over {layers} layers + lm_head")
for e in rows[:28]:
    print(f"{e.device_time_total/1e3:8.3f} ms  {e.count:4d}x  {e.key[:110]}")
