"""Host-side profile of `quantize_linear_` on the random-init Llama-3-8B (BASELINE configs[3]): wall time, device allocations
inside the call and the cProfile top entries -- the call is host-bound (GPU work: 22.7 GB at ~6 TB/s = 4 ms)."""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tools import llama_bench

layers = int(sys.argv[1]) if len(sys.argv) > 1 else None
for trial in range(2):
    from transformers import LlamaConfig, LlamaForCausalLM
    import torchmx_b200  # noqa: F401
    from torchmx_b200.config import MXConfig, QLinearConfig
    from torchmx_b200.quant_api import quantize_linear_
    kw = dict(llama_bench.SHAPES["8b"])
    if layers:
        kw["num_hidden_layers"] = layers
    cfg = LlamaConfig(max_position_embeddings=8192, **kw)
    torch.set_default_dtype(torch.bfloat16)
    with torch.device("cuda"):
        model = LlamaForCausalLM(cfg).eval()
    torch.set_default_dtype(torch.float32)
    qc = QLinearConfig(weights_config=MXConfig("float6_e3m2", 32), activations_config=MXConfig("float8_e4m3", 32))
    torch.cuda.synchronize()
    n0 = torch.cuda.memory_stats().get("num_device_alloc", 0)
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    if trial == 1:
        pr.enable()
    quantize_linear_(model, qc)
    if trial == 1:
        pr.disable()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f"trial {trial}: host issue {t_issue * 1e3:.1f} ms, with sync {t_all * 1e3:.1f} ms, cudaMalloc calls {torch.cuda.memory_stats().get('num_device_alloc', 0) - n0}", flush=True)
    if trial == 1:
        st = pstats.Stats(pr)
        st.sort_stats("tottime")
        buf = io.StringIO()
        st.stream = buf
        st.print_stats(25)
        print("\n".join(buf.getvalue().splitlines()[6:40]))
    del model
    torch.cuda.empty_cache()
