"""World-size-2 gloo tests (CPU) of the multi-GPU host logic in gpu_olap_engine_b200/sharded.py: destination routing,
count exchange, variable-size all-to-all, local-join hand-off and id translation.  The device kernels are replaced by a
numpy/oracle stand-in (test infrastructure); the GPU versions of the same operators are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O


class HostOps:
    """Stand-in for sharded.DeviceOps on CPU tensors (same contracts)."""

    def __init__(self):
        import gpu_olap_engine_b200 as g
        self.lib = g.load()  # olap_shuffle_dest_of is a host function of the C-ABI library: the routing is the product's

    def empty(self, n, dtype):
        return torch.empty(max(int(n), 0), dtype=dtype)

    def shuffle_partition(self, keys, n, bits, tag=""):
        """Same contract as DeviceOps.shuffle_partition: regions start on multiples of 16 rows, (begin, count) per rank."""
        k = keys[:n].numpy()
        dest = np.array([self.lib.olap_shuffle_dest_of(int(x), bits) for x in k], np.int64) if bits else np.zeros(n, np.int64)
        P = 1 << bits
        count = np.bincount(dest, minlength=P).tolist()
        begin, end = [], 0
        for c in count:
            b = (end + 15) // 16 * 16; begin.append(b); end = b + c
        ok = np.zeros(n + 16 * P, np.int64); oi = np.zeros(n + 16 * P, np.int32)
        for p in range(P):
            rows = np.nonzero(dest == p)[0]
            ok[begin[p]:begin[p] + count[p]] = k[rows]; oi[begin[p]:begin[p] + count[p]] = rows
        return torch.from_numpy(ok), torch.from_numpy(oi), begin + [end], count

    def hash_join(self, bk, nb, pk, npb, max_matches, strategy=-1):
        cnt, pairs = O.hash_join_i64(bk[:nb].numpy(), pk[:npb].numpy(), max_matches=max_matches)
        out = torch.zeros(2 * max(max_matches, 1), dtype=torch.int32)
        out[:2 * len(pairs)] = torch.from_numpy(pairs.astype(np.int32).reshape(-1))
        return cnt, out

    def group_by(self, keys, n, cols, funcs, types, hint):
        if keys is not None and n == 0:
            return 0, self.empty(0, torch.int64), [self.empty(0, torch.int64) for _ in funcs]
        vals = [(c[:n].numpy().view(np.float64) if t == 1 else c[:n].numpy()) if c is not None else np.zeros(n, np.int64) for c, t in zip(cols, types)]
        gk, res = O.group_by_i64(None if keys is None else keys[:n].numpy(), vals, funcs)
        return len(gk), torch.from_numpy(gk), [torch.from_numpy(r.view(np.int64).copy()) for r in res]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gpu_olap_engine_b200 import sharded
        ops = HostOps()
        # ---- join: each rank owns a contiguous slice of both sides
        nb, npb, dom = 3000, 7000, 5000
        gb = O.gen_uniform_i64(nb * world, 11, dom); gp = O.gen_uniform_i64(npb * world, 12, dom)
        sj = sharded.ShardedJoin(ops)
        res = sj.join(torch.from_numpy(gb[rank * nb:(rank + 1) * nb].copy()), nb, torch.from_numpy(gp[rank * npb:(rank + 1) * npb].copy()), npb,
                      max_matches=200000)
        pairs = sharded.ShardedJoin.global_pairs(res, [r * nb for r in range(world)], [r * npb for r in range(world)]).numpy()
        # every key lives on exactly one rank
        dest = {ops.lib.olap_shuffle_dest_of(int(k), sj.x.bits) for k in gb[pairs[:, 0]]} if len(pairs) else set()
        assert dest <= {rank}
        gathered = [None] * world
        dist.all_gather_object(gathered, pairs)
        # ---- exchange_regions(skip_self, prefix): the own region stays in place, the received rows of every column start at the
        # same 4-row-aligned base, column 1 carries the own ids in front (what the overlapped join indexes its pairs against)
        pk, pid, begin, count = ops.shuffle_partition(torch.from_numpy(gp[rank * npb:(rank + 1) * npb].copy()), npb, sj.x.bits)
        own = pid[begin[rank]:begin[rank] + count[rank]]
        (rk, ri), segs = sj.x.exchange_regions([pk, pid], begin, count, skip_self=True, prefix=[None, own])
        base = (count[rank] + 3) & ~3
        assert segs[0] == (rank, base) and sum(c for _, c in segs[1:]) == rk.numel() - base == ri.numel() - base
        assert (ri[:count[rank]] == own).all()
        peer = 1 - rank
        peer_rows = gp[peer * npb:(peer + 1) * npb]
        got_keys = rk[base:].numpy(); got_ids = ri[base:].numpy()
        assert all(src != rank or c == 0 for src, c in segs[1:])                 # nothing was sent to oneself
        assert (peer_rows[got_ids] == got_keys).all()                            # every received (key, id) is a row of the peer
        assert all(ops.lib.olap_shuffle_dest_of(int(k), sj.x.bits) == rank for k in got_keys)
        assert len(got_keys) == sum(1 for k in peer_rows if ops.lib.olap_shuffle_dest_of(int(k), sj.x.bits) == rank)
        # ---- group-by with AVG / MIN / MAX / COUNT over sharded rows
        n = 5000
        k = O.gen_uniform_i64(n * world, 4, 300); v = O.gen_uniform_i64(n * world, 5, 1 << 20); f = O.gen_dyadic_f64(n * world, 6, 1 << 20)
        sl = slice(rank * n, (rank + 1) * n)
        sg = sharded.ShardedGroupBy(ops)
        funcs = [O.SUM, O.COUNT, O.MIN, O.MAX, O.AVG, O.SUM, O.AVG]
        cols = [torch.from_numpy(v[sl].copy())] * 5 + [torch.from_numpy(f[sl].copy().view(np.int64))] * 2
        m, mk, out = sg.group_by(torch.from_numpy(k[sl].copy()), n, cols, funcs, [0] * 5 + [1] * 2, 300)
        shard = (mk.numpy(), [o.numpy() for o in out])
        gshards = [None] * world
        dist.all_gather_object(gshards, shard)
        if rank == 0:
            allp = np.concatenate(gathered)
            ec, ep = O.hash_join_i64(gb, gp)
            assert len(allp) == ec and (O.sort_pairs(allp.astype(np.uint32)) == ep).all()
            keys = np.concatenate([s[0] for s in gshards]); order = np.argsort(keys)
            ek, eres = O.group_by_i64(k, [v] * 5 + [f] * 2, funcs)
            assert len(np.unique(keys)) == len(keys) and (keys[order] == ek).all()       # shards are disjoint and complete
            for i, exp in enumerate(eres):
                got = np.concatenate([s[1][i] for s in gshards])[order]
                assert (got == exp).all(), i
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_sharded_join_and_group_by_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in results), results
