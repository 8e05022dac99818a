"""Key-range-sharded hash join across the GPUs of one box: ctypes callers of the olap_shard_* group of include/olap_b200.h.

Two drivers, same kernels:
  ShardGroup        one process, all ranks (olap_shard_ctx_create_all + olap_sharded_hash_join_i64).  The devices may repeat, so
                    the exchange runs on a 1-GPU box too (tests/test_gpu_shard.py).
  RangeShardedJoin  one process per GPU (torchrun): the slabs are torch symmetric memory (peer-mapped over NVSwitch), the
                    barriers between the phases are one-element NCCL all-reduces on the scatter stream.  torch is plumbing
                    here: device memory the peers can address, streams, the tiny collectives.
The reference has no multi-GPU path (README.md:414); this follows SURVEY.md 8(e).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .core import Context

HOT_SLOTS = 8192      # open-addressing slots of the heavy-hitter table (csrc/djoin.cu kHotSlots)
HOT_MAX_KEYS = HOT_SLOTS // 4   # at most a quarter of them hold a key (kHotMaxKeys): short probe chains for the rows that are NOT hot
FLAG_NAMES = ["region overflow (skewed keys)", "build key outside the planned range", "repeated build key", "-"]


class ShardGroup:
    """All ranks of a sharded join in one process; rank r lives on devices[r]."""

    def __init__(self, devices, build_rows_cap: int, probe_rows_cap: int, probe_chunks: int = 1, max_gpu_memory: int = 0):
        self.lib = _lib.load()
        self.n = len(devices)
        dev = (C.c_int32 * self.n)(*devices)
        ctxs = (C.c_void_p * self.n)(); shards = (C.c_void_p * self.n)()
        _lib.check(self.lib.olap_shard_ctx_create_all(self.n, dev, max_gpu_memory, build_rows_cap, probe_rows_cap, probe_chunks, ctxs, shards))
        self.shards = [shards[r] for r in range(self.n)]
        self.ctxs = []
        for r in range(self.n):
            c = object.__new__(Context)
            c.lib, c.handle, c.device = self.lib, ctxs[r], devices[r]
            self.ctxs.append(c)

    def close(self):
        for s in getattr(self, "shards", []):
            self.lib.olap_shard_destroy(s)
        self.shards = []
        for c in getattr(self, "ctxs", []):
            c.close()
        self.ctxs = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def join(self, build_keys, n_build, probe_keys, n_probe, out_pairs, max_matches, hot_pairs=None, max_hot_pairs=None):
        """build_keys[r] / probe_keys[r]: DeviceBuffers of rank r; out_pairs[r]: DeviceBuffer of 2 * max_matches[r] u32.
        hot_pairs[r] (optional): DeviceBuffer for rank r's {hot slot, local probe row} pairs -- turns on heavy-hitter handling.
        Returns the per-rank pair counts (and the per-rank hot pair counts when hot_pairs is given); raises
        OlapError(UNSUPPORTED) when the key domain is not dense / a flag was raised."""
        n = self.n
        arr = lambda xs: (C.c_void_p * n)(*[x.ptr if x is not None else None for x in xs])
        u64s = lambda xs: (C.c_uint64 * n)(*[int(x) for x in xs])
        counts = (C.c_uint64 * n)(); hot_counts = (C.c_uint64 * n)()
        sh = (C.c_void_p * n)(*self.shards)
        hot = hot_pairs is not None
        _lib.check(self.lib.olap_sharded_hash_join_i64(sh, n, arr(build_keys), u64s(n_build), arr(probe_keys), u64s(n_probe), arr(out_pairs),
                                                       u64s(max_matches), counts, arr(hot_pairs) if hot else None,
                                                       u64s(max_hot_pairs) if hot else None, hot_counts if hot else None))
        if hot:
            return [int(c) for c in counts], [int(c) for c in hot_counts]
        return [int(c) for c in counts]

    def group_by(self, keys, n_rows, aggs, n_groups_hint: int, out_capacity=None):
        """GROUP BY over all ranks' rows (olap_sharded_group_by_i64): keys[r] = rank r's key column, aggs[r] = [(func, values, value_type)]
        with the same functions / types on every rank.  Returns per rank (n_groups, keys ndarray, [typed ndarray per aggregate]) -- rank r
        holds the groups whose key it owns (olap_shuffle_dest_of); the union over the ranks is the result."""
        from .core import typed_agg_results
        n = self.n
        na = len(aggs[0])
        caps = [int(c) for c in (out_capacity if out_capacity is not None else [max(int(n_groups_hint), 1)] * n)]
        spec_arrays = [(_lib.AggSpec * na)(*[_lib.AggSpec(v.ptr if v is not None else None, int(f), int(t)) for f, v, t in aggs[r]]) for r in range(n)]
        ok = [self.ctxs[r].alloc(max(caps[r], 1), np.int64) for r in range(n)]
        oa = [self.ctxs[r].alloc(na * max(caps[r], 1), np.uint64) for r in range(n)]
        ctxs = (C.c_void_p * n)(*[c.handle for c in self.ctxs])
        kp = (C.c_void_p * n)(*[k.ptr if k is not None else None for k in keys])
        sp = (C.c_void_p * n)(*[C.addressof(a) for a in spec_arrays])
        okp = (C.c_void_p * n)(*[b.ptr for b in ok]); oap = (C.c_void_p * n)(*[b.ptr for b in oa])
        g = (C.c_uint64 * n)()
        _lib.check(self.lib.olap_sharded_group_by_i64(ctxs, n, kp, sp, na, (C.c_uint64 * n)(*[int(x) for x in n_rows]), int(n_groups_hint), okp, oap,
                                                      (C.c_uint64 * n)(*caps), g))
        out = []
        for r in range(n):
            ng = int(g[r])
            out.append((ng, ok[r].to_numpy(ng), typed_agg_results(oa[r].to_numpy(), max(caps[r], 1), ng, aggs[r])))
        return out

    def hot_pairs_to_global(self, r, hot_pairs, n_pairs, build_row_offsets, probe_row_offsets):
        """Global (left, right) row ids of rank r's heavy-hitter pairs (numpy uint64)."""
        tab = (C.c_uint64 * HOT_SLOTS)()
        _lib.check(self.lib.olap_shard_hot_table(self.shards[r], tab, None))
        tab = np.ctypeslib.as_array(tab).astype(np.uint64)
        pr = hot_pairs.to_numpy(2 * n_pairs).reshape(-1, 2).astype(np.uint64)
        v = tab[pr[:, 0]] - np.uint64(1)
        left = np.asarray(build_row_offsets, np.uint64)[(v >> np.uint64(32)).astype(np.int64)] + (v & np.uint64(0xFFFFFFFF))
        right = np.uint64(probe_row_offsets[r]) + pr[:, 1]
        return left, right

    def pairs_to_global(self, r, pairs, n_pairs, build_row_offsets, probe_row_offsets):
        """(left, right) global row ids of rank r's pairs, as two numpy uint64 arrays."""
        ctx = self.ctxs[r]
        bo = ctx.upload_raw(np.asarray(build_row_offsets, np.uint64)); po = ctx.upload_raw(np.asarray(probe_row_offsets, np.uint64))
        left = ctx.alloc(max(n_pairs, 1), np.uint64); right = ctx.alloc(max(n_pairs, 1), np.uint64)
        _lib.check(self.lib.olap_shard_pairs_to_global(self.shards[r], pairs.ptr, n_pairs, bo.ptr, po.ptr, left.ptr, right.ptr, None))
        return left.to_numpy(n_pairs), right.to_numpy(n_pairs)


class RangeShardedJoin:
    """One rank of the range-sharded join under torchrun (one process per GPU, NCCL process group initialised)."""

    def __init__(self, ctx: Context, build_rows_cap: int, probe_rows_cap: int, probe_chunks: int = 4, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.torch, self.dist = torch, dist
        self.ctx, self.lib = ctx, ctx.lib
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.chunks = probe_chunks
        self.device = torch.device("cuda", ctx.device)
        h = C.c_void_p()
        _lib.check(self.lib.olap_shard_create(ctx.handle, self.rank, self.world, build_rows_cap, probe_rows_cap, probe_chunks, C.byref(h)))
        self.handle = h.value
        self.chunk_cap = (probe_rows_cap + probe_chunks - 1) // probe_chunks
        nbytes = int(self.lib.olap_shard_slab_bytes(self.handle))
        self.slab = symm.empty(nbytes, dtype=torch.uint8, device=self.device)      # peer-mapped on every rank after the rendezvous
        self.symm_handle = symm.rendezvous(self.slab, self.group)
        ptrs = (C.c_void_p * self.world)(*[int(p) for p in self.symm_handle.buffer_ptrs])
        _lib.check(self.lib.olap_shard_attach(self.handle, ptrs))
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)          # every rank's arrival words are zeroed before anyone signals
        self.nccl_barrier = os.environ.get("OLAP_SHARD_NCCL_BARRIER", "0") == "1"
        self.comm = torch.cuda.ExternalStream(self.lib.olap_ctx_stream(ctx.handle, 1), device=self.device)   # peer copies / counts + barriers
        self.comp = torch.cuda.ExternalStream(self.lib.olap_ctx_stream(ctx.handle, 0), device=self.device)   # local partition + join
        self.ce_mode = not os.environ.get("OLAP_SHARD_MODE", "ce").lower().startswith("t")
        # ce mode: the scatters' partition kernels get a stream of their own, so that they are not queued behind a join that waits
        # for a barrier; tma mode: the shuffle kernel is the transfer, it runs on comm
        self.part = torch.cuda.ExternalStream(self.lib.olap_ctx_stream(ctx.handle, 2), device=self.device) if self.ce_mode else self.comm
        self._one = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._stats = {}
        self._done = None       # event: this rank's previous join has finished reading its slabs
        self._hot_glob = None   # merged heavy-hitter table of the last join: slot -> (source rank << 32 | source row) + 1

    def close(self):
        if getattr(self, "handle", None):
            self.lib.olap_shard_destroy(self.handle)
            self.handle = None

    def _barrier(self):
        """Device-ordered on the scatter stream: completes when every rank's work queued before it on that stream has.
        olap_shard_barrier (one warp, signals through the slabs); OLAP_SHARD_NCCL_BARRIER=1 uses a one-element all-reduce instead --
        which cannot start while a join CTA holds the SM's registers and so stalls the shuffle stream for the length of a join."""
        if self.nccl_barrier:
            with self.torch.cuda.stream(self.comm):
                self.dist.all_reduce(self._one, group=self.group)
        else:
            _lib.check(self.lib.olap_shard_barrier(self.handle, self.comm.cuda_stream))

    def _phase_done(self):
        """After a scatter + send on comm: let the peers know, and make comp wait for all of them -- comm itself goes on."""
        if self.nccl_barrier:
            self._barrier()
            ev = self.torch.cuda.Event(); ev.record(self.comm); self.comp.wait_event(ev)
        else:
            _lib.check(self.lib.olap_shard_signal(self.handle, self.comm.cuda_stream))
            _lib.check(self.lib.olap_shard_wait(self.handle, self.comp.cuda_stream))

    def _scatter(self, side, chunk, ptr, n, row_base):
        """Partition kernel on `part`, then -- ordered behind it -- the peer copies (ce mode) and the count publication on `comm`."""
        torch, lib, h = self.torch, self.lib, self.handle
        _lib.check(lib.olap_shard_scatter(h, side, chunk, ptr if n else None, n, row_base, self.part.cuda_stream))
        if self.part is not self.comm:
            ev = torch.cuda.Event(); ev.record(self.part); self.comm.wait_event(ev)
        _lib.check(lib.olap_shard_send(h, side, chunk, self.comm.cuda_stream))

    def _global_range(self, build, nb):
        key = (build.data_ptr(), nb)
        if key not in self._stats:
            torch, dist = self.torch, self.dist
            mn, mx = C.c_int64(), C.c_int64()
            _lib.check(self.lib.olap_shard_key_range(self.handle, build.data_ptr() if nb else None, nb, C.byref(mn), C.byref(mx), self.comm.cuda_stream))
            with torch.cuda.stream(self.comm):
                t = torch.tensor([-mn.value, mx.value], dtype=torch.int64, device=self.device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
                n = torch.tensor([nb], dtype=torch.int64, device=self.device)
                dist.all_reduce(n, group=self.group)
                t, n = t.tolist(), int(n.item())
            if len(self._stats) > 64:
                self._stats.clear()
            self._stats[key] = (-t[0], t[1], n)
        return self._stats[key]

    def _hot_keys(self, probe, npb, stream):
        """The agreed heavy-hitter set of this probe column: every rank samples its shard, the lists are all-gathered, the counts
        summed per key and the <= 2048 most frequent kept (same result on every rank).  Remembered per (pointer, rows)."""
        key = ("hot", probe.data_ptr(), npb)
        if key not in self._stats:
            torch, dist = self.torch, self.dist
            cap = 16384
            hk = (C.c_uint32 * cap)(); hc = (C.c_uint32 * cap)(); hn = C.c_uint32()
            _lib.check(self.lib.olap_shard_hot_candidates(self.handle, probe.data_ptr() if npb else None, npb, hk, hc, cap, C.byref(hn), stream))
            mine = torch.zeros(2 * cap + 1, dtype=torch.int64, device=self.device)
            n = int(hn.value)
            mine[0] = n
            if n:
                mine[1:1 + n] = torch.from_numpy(np.ctypeslib.as_array(hk)[:n].astype(np.int64)).to(self.device)
                mine[1 + cap:1 + cap + n] = torch.from_numpy(np.ctypeslib.as_array(hc)[:n].astype(np.int64)).to(self.device)
            allc = torch.empty(self.world * (2 * cap + 1), dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(allc, mine, group=self.group)
            allc = allc.view(self.world, -1).cpu().numpy()
            freq = {}
            for r in range(self.world):
                m = int(allc[r, 0])
                for k, c in zip(allc[r, 1:1 + m], allc[r, 1 + cap:1 + cap + m]):
                    freq[int(k)] = freq.get(int(k), 0) + int(c)
            top = sorted(freq.items(), key=lambda kv: (-kv[1], kv[0]))[:HOT_MAX_KEYS]
            # Is the heavy-hitter table worth its lookups?  A region has 25 % of room (~2^17 rows per 2^27-row chunk): only a key that
            # carries more than ~0.02 % of the probe rows can overflow it by itself.  Below that (Zipf s <~ 0.7 over 2^29 keys) every row
            # would pay the lookup for nothing -- measured at 2 GPUs, s = 0.5: 25.0 ms per step with the table, against 15.8 ms uniform.
            samples = self.world * (1 << 18)
            heaviest = top[0][1] if top else 0
            self._stats[key] = (np.array(sorted(k for k, _ in top), dtype=np.uint32), heaviest * 4096 >= samples)
        return self._stats[key]

    def join(self, build, nb, probe, npb, pairs, max_matches, hot=0, hot_pairs=None, max_hot_pairs=0):
        """build / probe: int64 CUDA tensors (this rank's shards); pairs: int32 tensor of 2 * max_matches.  hot = 1: heavy-hitter
        handling for Zipf-skewed probe keys (hot_pairs: int32 tensor of 2 * max_hot_pairs for the {hot slot, local probe row} pairs).
        Returns {"count", "hot_count", "flags" (OR over all ranks; any set = result invalid), "supported"}; everything is queued
        asynchronously and this returns after the rank's own work has finished."""
        torch, dist, lib, h = self.torch, self.dist, self.lib, self.handle
        gmin, gmax, gn = self._global_range(build, nb)
        hot_keys = None
        self.comm.wait_stream(torch.cuda.current_stream(self.device))     # the inputs were produced on the caller's stream
        if hot:   # the agreed heavy-hitter set (remembered per probe column) decides between the table (1) and room only (2)
            st = lib.olap_shard_plan(h, gmin, gmax, gn, 2, self.comm.cuda_stream)    # olap_shard_hot_candidates needs the planned geometry
            if st == _lib.ERR_UNSUPPORTED:
                return {"supported": False, "count": 0, "hot_count": 0, "flags": [0, 0, 0, 0]}
            _lib.check(st)
            hot_keys, needed = self._hot_keys(probe, npb, self.comm.cuda_stream)
            hot = 1 if needed else 2
        st = lib.olap_shard_plan(h, gmin, gmax, gn, hot, self.comm.cuda_stream)
        if st == _lib.ERR_UNSUPPORTED:
            return {"supported": False, "count": 0, "hot_count": 0, "flags": [0, 0, 0, 0]}
        _lib.check(st)
        comm, comp, part = self.comm, self.comp, self.part
        trace = [] if os.environ.get("OLAP_SHARD_TRACE") else None

        def mark(name, stream):
            if trace is not None:
                e = torch.cuda.Event(enable_timing=True); e.record(stream); trace.append((name, e))

        if self._done is not None:
            comm.wait_event(self._done)
        mark("start", comm)
        comm.wait_stream(torch.cuda.current_stream(self.device))     # the inputs were produced on the caller's stream
        if hot == 1:
            _lib.check(lib.olap_shard_hot_set(h, hot_keys.ctypes.data_as(C.POINTER(C.c_uint32)), len(hot_keys), comm.cuda_stream))
        self._barrier()                                              # every rank has finished reading what the previous join delivered
        if part is not comm:
            ev = torch.cuda.Event(); ev.record(comm); part.wait_event(ev)
        mark("comm: entry barrier done", comm)
        self._scatter(0, 0, build.data_ptr(), nb, 0)
        mark("comm: build sent", comm)
        self._phase_done()
        _lib.check(lib.olap_shard_build(h, comp.cuda_stream))
        mark("comp: build partitioned", comp)
        if hot == 1:   # the hot keys' build rows: resolved by their owners, MAX-reduced, back to every rank before the probe side moves
            glob = torch.empty(HOT_SLOTS, dtype=torch.int64, device=self.device)
            _lib.check(lib.olap_shard_hot_resolve(h, glob.data_ptr(), comp.cuda_stream))
            ev = torch.cuda.Event(); ev.record(comp); part.wait_event(ev)
            with torch.cuda.stream(part):
                dist.all_reduce(glob, op=dist.ReduceOp.MAX, group=self.group)
            _lib.check(lib.olap_shard_hot_finalize(h, glob.data_ptr(), hot_pairs.data_ptr() if hot_pairs is not None else None, max_hot_pairs,
                                                   part.cuda_stream))
            self._hot_glob = glob
        cc = self.chunk_cap
        for k in range(self.chunks):
            lo = min(k * cc, npb); hi = min(lo + cc, npb)
            self._scatter(1, k, probe.data_ptr() + 8 * lo, hi - lo, lo)
            mark(f"comm: chunk {k} sent", comm)
            self._phase_done()
            _lib.check(lib.olap_shard_probe(h, k, pairs.data_ptr(), max_matches, comp.cuda_stream))
            mark(f"comp: chunk {k} done", comp)
        cnt = C.c_uint64(); hcnt = C.c_uint64(); fl = (C.c_uint32 * 4)()
        comp.wait_stream(comm)
        comp.wait_stream(part)
        _lib.check(lib.olap_shard_finish(h, C.byref(cnt), C.byref(hcnt), fl, comp.cuda_stream))
        self._done = torch.cuda.Event(); self._done.record(comp)
        if trace is not None and self.rank == 0:
            torch.cuda.synchronize()
            t0 = trace[0][1]
            print("[shard trace] " + " | ".join(f"{n} @{t0.elapsed_time(e):.2f}" for n, e in trace[1:]), flush=True)
        with torch.cuda.stream(comm):
            f = torch.tensor(list(fl), dtype=torch.int32, device=self.device)
            dist.all_reduce(f, op=dist.ReduceOp.MAX, group=self.group)
            flags = f.tolist()
        return {"supported": True, "count": int(cnt.value), "hot_count": int(hcnt.value), "flags": flags}

    def pairs_to_global(self, pairs, n_pairs, build_row_offsets, probe_row_offsets):
        """Global (left, right) row ids of this rank's pairs as two int64 CUDA tensors."""
        torch = self.torch
        bo = torch.tensor(list(build_row_offsets), dtype=torch.int64, device=self.device)
        po = torch.tensor(list(probe_row_offsets), dtype=torch.int64, device=self.device)
        left = torch.empty(max(n_pairs, 1), dtype=torch.int64, device=self.device); right = torch.empty_like(left)
        torch.cuda.current_stream(self.device).synchronize()
        _lib.check(self.lib.olap_shard_pairs_to_global(self.handle, pairs.data_ptr(), n_pairs, bo.data_ptr(), po.data_ptr(), left.data_ptr(),
                                                       right.data_ptr(), self.comp.cuda_stream))
        self.comp.synchronize()
        return left[:n_pairs], right[:n_pairs]
