"""torchrun entry (one process per GPU, NCCL): parity of the key-sharded join and group-by against the oracle.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/run_sharded_nccl.py
Used by tests/test_gpu_sharded.py when at least two GPUs are visible, and directly under `gpurun --gpus N`."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_olap_engine_b200 as g  # noqa: E402
import oracle as O  # noqa: E402
from gpu_olap_engine_b200 import sharded  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = g.Context(local, 0, 12)
    ops = sharded.DeviceOps(ctx)
    torch.cuda.set_stream(torch.cuda.ExternalStream(ctx.lib.olap_ctx_stream(ctx.handle, 0), device=ops.device))
    dev = ops.device
    for nb, npb, dom in [(200000, 900000, 150000), (1 << 21, 1 << 22, 1 << 22), (1000, 0, 10)]:
        gb = O.gen_uniform_i64(nb * world, 11, dom); gp = O.gen_uniform_i64(npb * world, 12, dom)
        sj = sharded.ShardedJoin(ops)
        res = sj.join(torch.from_numpy(gb[rank * nb:(rank + 1) * nb].copy()).to(dev), nb,
                      torch.from_numpy(gp[rank * npb:(rank + 1) * npb].copy()).to(dev), npb, max_matches=max(4 * npb * world, 16))   # matches per probe row grow with the world size (fixed key domain)
        pairs = sharded.ShardedJoin.global_pairs(res, [r * nb for r in range(world)], [r * npb for r in range(world)]).cpu().numpy()
        gathered = [None] * world
        dist.all_gather_object(gathered, pairs)
        if rank == 0:
            allp = np.concatenate(gathered)
            ec, ep = O.hash_join_i64(gb, gp, threads=4)
            assert len(allp) == ec, (len(allp), ec)
            assert (O.sort_pairs(allp.astype(np.uint32)) == O.sort_pairs(ep)).all()
            print(f"sharded join ok: world={world} {nb}x{npb} per rank, {ec} pairs")
    # the range-sharded path (dense global key domain): peer-store shuffle into fixed slabs, chunked probe side
    from gpu_olap_engine_b200.shard import RangeShardedJoin
    for nb, npb, chunks in [((1 << 20) + 3, (1 << 22) + 11, 4), (1 << 18, 1 << 18, 1)]:
        rng = np.random.default_rng(99)
        perm = rng.permutation(nb * world).astype(np.int64) + 77
        gp = rng.integers(0, nb * world + 200, npb * world).astype(np.int64)
        rj = RangeShardedJoin(ctx, nb, npb, chunks)
        b = torch.from_numpy(perm[rank * nb:(rank + 1) * nb].copy()).to(dev); p = torch.from_numpy(gp[rank * npb:(rank + 1) * npb].copy()).to(dev)
        mm = 2 * npb + 4096
        pairs = torch.empty(2 * mm, dtype=torch.int32, device=dev)
        for rep in range(3):
            res = rj.join(b, nb, p, npb, pairs, mm)
            assert res["supported"] and not any(res["flags"]), res
        left, right = rj.pairs_to_global(pairs, res["count"], [r * nb for r in range(world)], [r * npb for r in range(world)])
        mine = torch.stack([left, right], 1).cpu().numpy()
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            allp = np.concatenate(gathered)
            ec, ep = O.hash_join_i64(perm, gp, threads=4)
            assert len(allp) == ec, (len(allp), ec)
            assert (O.sort_pairs(allp.astype(np.uint32)) == O.sort_pairs(ep)).all()
            print(f"range-sharded join ok: world={world} {nb}x{npb} per rank, {chunks} chunks, {ec} pairs")
        rj.close()
    # Zipf-skewed probe keys: heavy hitters agreed on by all ranks, their probe rows joined at the source rank
    nb, npb = 1 << 19, 1 << 21
    rng = np.random.default_rng(5)
    perm = rng.permutation(nb * world).astype(np.int64) + 9
    w = np.arange(1, nb * world + 1, dtype=np.float64) ** -1.2
    cdf = np.cumsum(w); cdf /= cdf[-1]
    gp = perm[np.minimum(np.searchsorted(cdf, rng.random(npb * world)), nb * world - 1)]
    rj = RangeShardedJoin(ctx, nb, npb, 2)
    b = torch.from_numpy(perm[rank * nb:(rank + 1) * nb].copy()).to(dev); p = torch.from_numpy(gp[rank * npb:(rank + 1) * npb].copy()).to(dev)
    mm = 4 * npb
    pairs = torch.empty(2 * mm, dtype=torch.int32, device=dev); hot = torch.empty(2 * npb, dtype=torch.int32, device=dev)
    for rep in range(2):
        res = rj.join(b, nb, p, npb, pairs, mm, hot=1, hot_pairs=hot, max_hot_pairs=npb)
        assert res["supported"] and not any(res["flags"]), res
    left, right = rj.pairs_to_global(pairs, res["count"], [r * nb for r in range(world)], [r * npb for r in range(world)])
    hp = hot[:2 * res["hot_count"]].view(-1, 2).to(torch.int64) & 0xFFFFFFFF
    v = rj._hot_glob[hp[:, 0]] - 1
    hl = (v >> 32) * nb + (v & 0xFFFFFFFF); hr = rank * npb + hp[:, 1]
    mine = torch.cat([torch.stack([left, right], 1), torch.stack([hl, hr], 1)]).cpu().numpy()
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, res["hot_count"]))
    if rank == 0:
        allp = np.concatenate([g_[0] for g_ in gathered])
        ec, ep = O.hash_join_i64(perm, gp, threads=4)
        assert len(allp) == ec, (len(allp), ec)
        assert (O.sort_pairs(allp.astype(np.uint32)) == O.sort_pairs(ep)).all()
        nhot = sum(g_[1] for g_ in gathered)
        assert nhot > 0.2 * ec, (nhot, ec)
        print(f"range-sharded zipf join ok: world={world}, {ec} pairs of which {nhot} joined at their source rank")
    rj.close()
    n = 500000
    for G in (300, 200000):
        k = O.gen_uniform_i64(n * world, 4, G); v = O.gen_uniform_i64(n * world, 5, 1 << 20); f = O.gen_dyadic_f64(n * world, 6, 1 << 20)
        sl = slice(rank * n, (rank + 1) * n)
        funcs = [O.SUM, O.COUNT, O.MIN, O.MAX, O.AVG, O.SUM, O.AVG]
        cols = [torch.from_numpy(v[sl].copy()).to(dev)] * 5 + [torch.from_numpy(f[sl].copy().view(np.int64)).to(dev)] * 2
        sg = sharded.ShardedGroupBy(ops)
        m, mk, out = sg.group_by(torch.from_numpy(k[sl].copy()).to(dev), n, cols, funcs, [0] * 5 + [1] * 2, G)
        shard = (mk.cpu().numpy(), [o.cpu().numpy() for o in out])
        gs = [None] * world
        dist.all_gather_object(gs, shard)
        if rank == 0:
            keys = np.concatenate([s[0] for s in gs]); order = np.argsort(keys)
            ek, eres = O.group_by_i64(k, [v] * 5 + [f] * 2, funcs)
            assert len(np.unique(keys)) == len(keys) and (keys[order] == ek).all()
            for i, exp in enumerate(eres):
                got = np.concatenate([s[1][i] for s in gs])[order]
                assert (got == exp).all(), i
            print(f"sharded group-by ok: world={world} G={G}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
