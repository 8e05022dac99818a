# Is torch symmetric memory usable on this box?  Remote-store bandwidth of a plain torch copy into the peer's buffer.
import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 1 << 28  # 2 GiB of int64
t = sm.empty(n, dtype=torch.int64, device=torch.device("cuda", lr))
h = sm.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok", h.rank, h.world_size, [hex(p) for p in h.buffer_ptrs], "multicast", h.has_multicast_support, flush=True)
peer = (rank + 1) % world
pb = h.get_buffer(peer, (n,), torch.int64)
src = torch.full((n // 2,), rank + 1, dtype=torch.int64, device="cuda")
h.barrier()
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pb[rank * (n // 2):(rank + 1) * (n // 2)].copy_(src)   # remote store into the peer's buffer
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(rank, f"peer copy {src.numel() * 8 / dt / 1e9:.1f} GB/s", flush=True)
h.barrier()
got = t[peer * (n // 2):(peer + 1) * (n // 2)]
assert bool((got == peer + 1).all()), "peer data not visible"
print(rank, "peer stores visible after barrier", flush=True)
dist.destroy_process_group()
