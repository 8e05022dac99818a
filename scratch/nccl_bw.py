import os, torch, torch.distributed as dist, time
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for nbytes in [64 << 20, 512 << 20, 2 << 30, 4 << 30]:
    n = nbytes // 8
    src = torch.empty(n, dtype=torch.int64, device="cuda"); dst = torch.empty_like(src)
    split = [n // world] * world
    for _ in range(3): dist.all_to_all_single(dst, src, split, split)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): dist.all_to_all_single(dst, src, split, split)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out_bytes = nbytes * (world - 1) / world
    if rank == 0: print(f"all_to_all_single {nbytes >> 20} MB per rank: {ms:.2f} ms, {out_bytes / ms / 1e6:.0f} GB/s out per rank", flush=True)
    # plain send/recv peer copy
    if world == 2:
        for _ in range(2): dst[: n // 2].copy_(src[: n // 2])
    del src, dst
dist.destroy_process_group()
