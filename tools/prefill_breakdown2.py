"""One eager 2048-token prefill of the HF Llama-3-8B shape (2 layers) quantized with quantize_llm_(fuse_rmsnorm=True), for an ncu
launch list."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_bench as lb
with torch.no_grad():
    model, cfg, info = lb.build("8b", 2, "float6_e3m2", "float8_e4m3", llm_api=True, fuse_norm=True)
    ids = torch.randint(0, cfg.vocab_size, (1, 2048), device="cuda")
    for i in range(2):
        model(input_ids=ids, use_cache=False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    model(input_ids=ids, use_cache=False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done")
