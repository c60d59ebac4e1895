"""Context for the output-bound attention bmm: cuBLAS bf16 bmm of the same shape and a plain 268 MB fill."""
import torch
q = torch.randn(32, 2048, 128, device="cuda", dtype=torch.bfloat16)
k = torch.randn(32, 2048, 128, device="cuda", dtype=torch.bfloat16)
out = torch.empty(32, 2048, 2048, device="cuda", dtype=torch.bfloat16)
def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(n): fn()
    ts = []
    for r in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n * 1e3)
    return min(ts[1:])
print("cuBLAS bf16 bmm 32x2048x2048x128: %.1f us" % timed(lambda: torch.bmm(q, k.transpose(1, 2), out=out)))
print("fill 268 MB: %.1f us" % timed(lambda: out.fill_(1.0)))
src = torch.randn(32, 2048, 2048, device="cuda", dtype=torch.bfloat16)
print("copy 268 MB -> 268 MB: %.1f us" % timed(lambda: out.copy_(src)))
