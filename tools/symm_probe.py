import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
t = symm_mem.empty((1024, 8192), dtype=torch.bfloat16, device=torch.device("cuda", local))
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
if rank == 0:
    print("attrs:", [a for a in dir(hdl) if not a.startswith("_")])
    print("multicast_ptr:", hex(hdl.multicast_ptr), "buffer_ptrs:", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs:", [hex(p) for p in hdl.signal_pad_ptrs])
    print("world", world, "buffer_size", hdl.buffer_size, "signal_pad_size", hdl.signal_pad_size)
t.fill_(rank + 1)
hdl.barrier(channel=0)
# read peer buffer through P2P
peer = hdl.get_buffer((rank + 1) % world, (4,), torch.bfloat16)
print(rank, "peer sample", peer.tolist())
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    hdl.barrier(channel=1)
torch.cuda.synchronize()
try:
    with torch.cuda.graph(g):
        hdl.barrier(channel=1)
    g.replay(); torch.cuda.synchronize()
    print(rank, "barrier captured in a CUDA graph OK")
except Exception as e:
    print(rank, "graph capture of barrier failed:", type(e).__name__, str(e)[:200])
dist.barrier(); dist.destroy_process_group()
