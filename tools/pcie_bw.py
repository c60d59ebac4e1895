"""Raw PCIe copy bandwidth of the box (pinned host memory, 512 MiB copies): each direction alone and both at once."""
import torch, time, json
n = 512 << 20
h1, h2 = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d1, d2 = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return round(reps * n / dt / 1e9, 1)
run(True, True, 2)
print(json.dumps({"h2d_alone_GBps": run(True, False), "d2h_alone_GBps": run(False, True), "each_direction_when_both_GBps": run(True, True)}))
