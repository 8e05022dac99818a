import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, '.')
import gpu_olap_engine_b200 as g
from gpu_olap_engine_b200 import sharded
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = g.Context(local, 0, 8)
ops = sharded.DeviceOps(ctx)
torch.cuda.set_stream(torch.cuda.ExternalStream(ctx.lib.olap_ctx_stream(ctx.handle, 0), device=ops.device))
sj = sharded.ShardedJoin(ops)
nb, npb = 1 << 27, 1 << 29
build = ops.empty(nb, torch.int64); probe = ops.empty(npb, torch.int64)
ctx.generate(nb, 1, 3, nb * world, row0=rank * nb, out=build.data_ptr()); ctx.generate(npb, 0, 33, nb * world, row0=rank * npb, out=probe.data_ptr())
def T(fn, name):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    if rank == 0: print(f"{name}: {dt:.2f} ms", flush=True)
    return r
for it in range(3):
    if rank == 0: print("--- iter", it)
    pk, pid, begin = T(lambda: ops.shuffle_partition(probe, npb, sj.x.bits, "_p"), "partition probe (incl. begin.tolist)")
    send = [begin[p + 1] - begin[p] for p in range(world)]
    recv = T(lambda: sj.x.exchange_counts(send), "exchange counts")
    ok = ops.buffer("recv_k_p", sum(recv), torch.int64); oi = ops.buffer("recv_i_p", sum(recv), torch.int32)
    T(lambda: sj.x.all_to_all(pk, send, recv, ok), "a2a keys (%.2f GB out)" % ((sum(send) - send[rank]) * 8 / 1e9))
    T(lambda: sj.x.all_to_all(pid, send, recv, oi), "a2a ids")
    T(lambda: sj.shuffle_both(build, nb, probe, npb), "shuffle_both")
    T(lambda: ops.hash_join(ok, int(ok.numel()), ok, 0, 16), "noop join")
dist.destroy_process_group()
