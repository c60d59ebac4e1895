"""Developer build (MXQ_DEV=1): per-CTA globaltimer stamps of back-to-back decode-sized MX linears in one CUDA graph."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchmx_b200  # noqa
from torchmx_b200 import dtypes, mx_gemm
from torchmx_b200.mx_tensor import MXTensor
wdt = getattr(dtypes, os.environ.get("GT_W", "float8_e4m3"))
for shape in os.environ.get("GT_SHAPES", "4096x4096,14336x4096").split(","):
    N, K = (int(v) for v in shape.split("x"))
    X = MXTensor.to_mx(torch.randn(32, K, device="cuda", dtype=torch.bfloat16), dtypes.float8_e4m3, 32)
    Ws = [MXTensor.to_mx(torch.randn(N, K, device="cuda", dtype=torch.bfloat16), wdt, 32) for _ in range(8)]
    for W in Ws:
        mx_gemm.mark_static(W); torch.nn.functional.linear(X, W)
    tr = torch.zeros(16 * 16384, dtype=torch.int64, device="cuda")
    os.environ["MXQ_SKINNY_TRACE"] = hex(tr.data_ptr())
    g, st = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for W in Ws:
                y = torch.nn.functional.linear(X, W)
    del os.environ["MXQ_SKINNY_TRACE"]
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    full = tr.view(16, 1024, 16).cpu()
    t, ck = full[:, :, :8], full[:, :, 8:]
    used = [i for i in range(16) if int(t[i, 0, 0]) != 0]
    print(f"== {shape} {wdt.name}: launches traced {used}")
    t0 = int(t[used[0], :, 0][t[used[0], :, 0] != 0].min())
    names = ["entry", "prologue", "pdl_wait", "first_full", "last_issue", "acc_done", "stored", "exit"]
    for i in used:
        n_cta = int((t[i, :, 0] != 0).sum())
        rows = t[i, :n_cta].double() - t0
        line = "  ".join(f"{names[j]} {rows[:, j].min():7.0f}/{rows[:, j].median():7.0f}/{rows[:, j].max():7.0f}" for j in range(8))
        print(f"launch {i} ({n_cta} CTAs) ns min/med/max: {line}")
        c = ck[i, :n_cta].double()
        d = c[:, 1:] - c[:, :-1]
        if os.environ.get("MXQ_SKINNY_EXP") == "32":
            print(f"     split-K tail, cycles (median): stored->sync1 {(c[:, 1] - c[:, 6]).median():.0f}  sync1->reduced {(c[:, 2] - c[:, 1]).median():.0f}  reduced->sync2 {(c[:, 3] - c[:, 2]).median():.0f}  sync2->exit {(c[:, 7] - c[:, 3]).median():.0f}   (max stored->sync1 {(c[:, 1] - c[:, 6]).max():.0f}, min {(c[:, 1] - c[:, 6]).min():.0f})")
        print("     cycles between stamps (median over CTAs): " + "  ".join(f"{names[j]}->{names[j+1]} {d[:, j].median():.0f}" for j in range(7)))
