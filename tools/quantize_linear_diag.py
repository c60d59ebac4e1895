"""quantize_linear_ on the HF Llama-3-8B shape: wall / GPU time and what the caching allocator did (how many cudaMalloc calls,
how many bytes newly reserved), per trial.  MXQ_DIAG_EXPANDABLE=1 runs with expandable segments."""
import os, sys, time, torch
if os.environ.get("MXQ_DIAG_EXPANDABLE") == "1":
    os.environ["PYTORCH_CUDA_ALLOC_CONF"] = "expandable_segments:True"
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama_bench as lb
import torchmx_b200  # noqa
from torchmx_b200.config import MXConfig, QLinearConfig
from torchmx_b200.quant_api import quantize_linear_
qc = QLinearConfig(weights_config=MXConfig("float6_e3m2", 32), activations_config=MXConfig("float8_e4m3", 32))
for trial in range(3):
    with torch.no_grad():
        model, cfg, info = lb.build("8b", None, "float6_e3m2", "float8_e4m3", quantize=False)
    torch.cuda.synchronize()
    s0 = torch.cuda.memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    quantize_linear_(model, qc)
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
    s1 = torch.cuda.memory_stats()
    print(f"trial {trial}: host issue {1e3*(t1-t0):.1f} ms, gpu {e0.elapsed_time(e1):.1f} ms, cudaMalloc calls {s1['num_device_alloc']-s0['num_device_alloc']}, "
          f"cudaFree calls {s1['num_device_free']-s0['num_device_free']}, reserved +{(s1['reserved_bytes.all.current']-s0['reserved_bytes.all.current'])/1e9:.2f} GB, "
          f"allocated {s0['allocated_bytes.all.current']/1e9:.2f} -> {s1['allocated_bytes.all.current']/1e9:.2f} GB, alloc retries {s1['num_alloc_retries']-s0['num_alloc_retries']}", flush=True)
    del model
    torch.cuda.empty_cache()

if os.environ.get("MXQ_DIAG_LLM") == "1":  # the same for quantize_llm_ (MX attention / MLP blocks, fused norms), with a host profile
    import cProfile, pstats, io
    from torchmx_b200.config import QAttentionConfig
    from torchmx_b200.quant_api import quantize_llm_
    for trial in range(3):
        with torch.no_grad():
            model, cfg, info = lb.build("8b", None, "float6_e3m2", "float8_e4m3", quantize=False)
        torch.cuda.synchronize()
        s0 = torch.cuda.memory_stats()
        pr = cProfile.Profile() if trial == 2 else None
        t0 = time.perf_counter()
        if pr: pr.enable()
        quantize_llm_(model, QAttentionConfig(projection_config=qc), qc, fuse_rmsnorm=True)
        if pr: pr.disable()
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        s1 = torch.cuda.memory_stats()
        print(f"llm trial {trial}: host issue {1e3*(t1-t0):.1f} ms, with sync {1e3*(t2-t0):.1f} ms, cudaMalloc calls {s1['num_device_alloc']-s0['num_device_alloc']}, "
              f"reserved +{(s1['reserved_bytes.all.current']-s0['reserved_bytes.all.current'])/1e9:.2f} GB", flush=True)
        if pr:
            st = pstats.Stats(pr); st.sort_stats("tottime"); buf = io.StringIO(); st.stream = buf; st.print_stats(12); print(buf.getvalue())
        del model
        torch.cuda.empty_cache()
