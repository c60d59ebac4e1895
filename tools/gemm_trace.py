"""Timeline of pair 0's leader CTA (clock64 stamps) for one MX GEMM launch: GT_SHAPES=MxNxK."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchmx_b200  # noqa
from torchmx_b200 import dtypes
from torchmx_b200.mx_tensor import MXTensor
for shape in os.environ.get("GT_SHAPES", "8192x8192x8192").split(","):
    M, N, K = (int(v) for v in shape.split("x"))
    batch = int(os.environ.get("GT_BATCH", "0"))
    lead = (batch,) if batch else ()
    A = MXTensor.to_mx(torch.randn(*lead, M, K, device="cuda", dtype=torch.bfloat16), dtypes.float8_e4m3, 32)
    B = MXTensor.to_mx(torch.randn(*lead, N, K, device="cuda", dtype=torch.bfloat16), dtypes.float6_e3m2, 32)
    run = (lambda: torch.bmm(A, B.transpose(1, 2))) if batch else (lambda: torch.nn.functional.linear(A, B))
    run()
    tr = torch.zeros(1024, dtype=torch.int64, device="cuda")  # [0,256): 8 per tile; [256,512): tile 2, 4 per k-block
    os.environ["MXQ_GEMM_TRACE"] = hex(tr.data_ptr())
    run()
    torch.cuda.synchronize()
    del os.environ["MXQ_GEMM_TRACE"]
    kbt = tr[256:512].view(64, 4).cpu()
    fin = tr[512:640].cpu(); st = tr[640:768].cpu()
    t = tr[:512].view(64, 8).cpu()[:32]
    t0 = int(t[0, 0])
    print(f"== {shape}: k_blocks={K//128}  (cycles since the MMA warp's first stamp)")
    print("tile | mma: wait_empty_start  empty_ok  first_full_ok  issued_all | epi: full_seen  released  done | gaps: issue->full_seen  full_seen->released  released->next_empty_ok")
    n = int((t[:, 0] != 0).sum())
    for i in range(n):
        r = [int(x) - t0 for x in t[i, :7]]
        nxt = int(t[i + 1, 1]) - t0 if i + 1 < n else None
        gaps = f"{r[4]-r[3]:6d} {r[5]-r[4]:6d} {(nxt - r[5]) if nxt is not None else -1:6d}"
        print(f"{i:3d} | {r[0]:8d} {r[1]:8d} {r[2]:8d} {r[3]:8d} | {r[4]:8d} {r[5]:8d} {r[6]:8d} | {gaps}")
    print("tile 2 per k-block: loop_top  sf_ok(+)  full_ok(+)")
    for kb in range(min(K // 128, 64)):
        a, b, c = (int(x) - t0 for x in kbt[kb, :3])
        print(f"  kb {kb:2d}: {a:8d}  +{b-a:5d}  +{c-b:5d}")
    npair = int((fin != 0).sum())
    base = int(st[:npair].min())
    import statistics
    starts = [int(x) - base for x in st[:npair]]; ends = [int(x) - base for x in fin[:npair]]
    tiles_total = ((M + 255) // 256) * ((N + 255) // 256) * max(batch, 1)
    print(f"pairs={npair} tiles={tiles_total}: start spread {max(starts)} ns; finish min {min(ends)} median {statistics.median(ends)} max {max(ends)} ns")
    print("finish by pair:", ends)
