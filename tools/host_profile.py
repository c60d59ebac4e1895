"""Where the host time of one layer goes (cProfile): from_float (weight quantization) and forward (eager)."""
import cProfile, pstats, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchmx_b200  # noqa
from torchmx_b200.config import MXConfig, QLinearConfig
from torchmx_b200.layers.mx_linear import MXInferenceLinear
qc = QLinearConfig(weights_config=MXConfig("float6_e3m2", 32), activations_config=MXConfig("float8_e4m3", 32))
lin = torch.nn.Linear(1024, 1024, bias=False).to("cuda", torch.bfloat16)
x = torch.randn(32, 1024, device="cuda", dtype=torch.bfloat16)
with torch.no_grad():
    m = MXInferenceLinear.from_float(lin, qc); m(x); torch.cuda.synchronize()
    for name, fn, n in (("from_float", lambda: MXInferenceLinear.from_float(lin, qc), 300), ("forward", lambda: m(x), 300)):
        t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize()
        print(f"{name}: {(time.perf_counter()-t0)/n*1e6:.1f} us per call (host-bound loop)")
        pr = cProfile.Profile(); pr.enable()
        for _ in range(n): fn()
        pr.disable(); torch.cuda.synchronize()
        st = pstats.Stats(pr); st.sort_stats("cumulative")
        import io; buf = io.StringIO(); st.stream = buf; st.print_stats(18); print("\n".join(buf.getvalue().splitlines()[6:30]))
