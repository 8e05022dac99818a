import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = (4 << 30) // 8
src = torch.empty(n + 64, dtype=torch.int64, device="cuda"); dst = torch.empty(n + 64, dtype=torch.int64, device="cuda")
def run(name, send, recv, soff=0, doff=0):
    s = src[soff:soff + sum(send)]; d = dst[doff:doff + sum(recv)]
    for _ in range(2): dist.all_to_all_single(d, s, recv, send)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4): dist.all_to_all_single(d, s, recv, send)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    if rank == 0: print(f"{name}: {ms:.2f} ms  {send[1 - rank] * 8 / ms / 1e6:.0f} GB/s out", flush=True)
h = n // 2
run("aligned halves", [h, h], [h, h])
run("split at +1 row (8 B misaligned both sides equally)", [h + 1, h - 1] if rank == 0 else [h + 1, h - 1], [h + 1, h + 1] if rank == 0 else [h - 1, h - 1])
run("send offset +1 row only", [h, h], [h, h], soff=1)
run("recv offset +1 row only", [h, h], [h, h], doff=1)
run("both offset +1 row", [h, h], [h, h], soff=1, doff=1)
dist.destroy_process_group()
