"""Key-sharded hash join and group-by across the GPUs of one box: one process per GPU, torch.distributed (NCCL over
NVLink 5 / NVSwitch) for the exchange, the sm_100a kernels behind the C ABI for everything else.

The reference has no multi-GPU path ("No multi-GPU support yet", README.md:414); this follows SURVEY.md 8(e):
  join      dest = top bits of mix64(key ^ salt)  ->  olap_shuffle_partition_i64 writes per-destination contiguous send
            regions (no staging copy)  ->  count exchange  ->  all-to-all of (key, source row id)  ->  independent local
            joins (identical to the 1-GPU path; the salt keeps the destination bits independent of the local radix bits)
  group-by  local pre-aggregation (cuts the volume to <= G tuples per GPU)  ->  all-to-all of (key, partials) by the
            same destination hash  ->  per-shard merge (SUM/COUNT add, MIN/MAX combine, AVG carried as (sum, count))
Output stays sharded.  Row ids are (source rank, local row id): pairs index the receive buffers, whose per-source
segments and shipped local ids give the global id (see global_pairs).

The device operations are injected (``ops``) so that the exchange logic runs unchanged under gloo on CPU in tests/.
"""
from __future__ import annotations

import ctypes as C
import json
import sys
import math
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .core import AVG, COUNT, F64, I64, JOIN_AUTO, MAX, MIN, SUM, AggSpec


def _log2(world: int) -> int:
    bits = int(math.log2(world))
    if (1 << bits) != world:
        raise ValueError("key sharding needs a power-of-two number of ranks")
    return bits


class DeviceOps:
    """GPU implementation of the operations the sharded operators need: C-ABI calls on torch CUDA tensors."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.lib = ctx.lib
        self.device = torch.device("cuda", ctx.device)

    def _stream(self):
        # torch's default stream is the legacy stream (handle 0); the C ABI reads NULL as "the context's stream 0",
        # so pass cudaStreamLegacy (0x1) explicitly to stay ordered with torch's own work
        return torch.cuda.current_stream(self.device).cuda_stream or 1

    def empty(self, n, dtype):
        return torch.empty(max(int(n), 0), dtype=dtype, device=self.device)

    def buffer(self, name, n, dtype):
        """Cached scratch tensor: multi-GB send/receive buffers are reused across calls instead of re-allocated."""
        cache = self.__dict__.setdefault("_buffers", {})
        t = cache.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(max(int(n * 1.05) + 1024, 1), dtype=dtype, device=self.device)
            cache[name] = t
        return t[:n]

    def shuffle_partition_async(self, keys, n, bits, tag=""):
        """Partition rows by destination rank: (keys, local row ids, meta) where meta = begin[2^bits + 1] ++ count[2^bits]
        is still on the device.  Region p = [begin[p], begin[p] + count[p]) starts on a multiple of 16 rows."""
        P = 1 << bits
        out_k = self.buffer("send_k" + tag, n + 16 * P, torch.int64); out_i = self.buffer("send_i" + tag, n + 16 * P, torch.int32)
        meta = self.empty(2 * P + 1, torch.int32)
        if bits == 0:
            out_k[:n].copy_(keys[:n]); out_i[:n].copy_(torch.arange(n, dtype=torch.int32, device=self.device))
            meta.copy_(torch.tensor([0, n, n], dtype=torch.int32))
            return out_k, out_i, meta
        _lib.check(self.lib.olap_shuffle_partition_i64(self.ctx.handle, keys.data_ptr(), None, n, bits, out_k.data_ptr(), out_i.data_ptr(),
                                                       meta.data_ptr(), meta.data_ptr() + 4 * (P + 1), self._stream()))
        return out_k, out_i, meta

    def shuffle_partition(self, keys, n, bits, tag=""):
        out_k, out_i, meta = self.shuffle_partition_async(keys, n, bits, tag)
        m = meta.tolist(); P = 1 << bits
        return out_k, out_i, m[:P + 1], m[P + 1:]

    def hash_join(self, build_keys, nb, probe_keys, npb, max_matches, strategy=JOIN_AUTO):
        pairs = self.buffer("pairs", 2 * max(max_matches, 1), torch.int32)
        cnt = C.c_uint64()
        st = self.lib.olap_hash_join_i64(self.ctx.handle, build_keys.data_ptr() if nb else None, nb, probe_keys.data_ptr() if npb else None, npb,
                                         pairs.data_ptr(), max_matches, None, C.byref(cnt), strategy, self._stream())
        _lib.check(st)
        return cnt.value, pairs

    def join_build(self, build_keys, nb, strategy=JOIN_AUTO):
        """Build side prepared once (olap_join_build_create); probe it batch by batch with join_probe."""
        from .core import JoinBuild
        return JoinBuild(self.ctx, build_keys.data_ptr() if nb else None, nb, strategy, stream=self._stream())

    def join_probe(self, jb, keys_ptr, n, row_base, pairs_ptr, max_matches):
        return jb.probe(keys_ptr if n else None, n, probe_row_base=row_base, out_pairs=pairs_ptr, max_matches=max_matches, stream=self._stream())

    def group_by(self, keys, n, cols, funcs, types, hint):
        """Returns (n_groups, keys[G], [col[G] as int64 bit patterns])."""
        na = len(funcs)
        cap = max(int(hint), 1)
        if keys is not None and n == 0:   # grouped aggregation of nothing: no groups (a NULL key pointer would mean "global")
            return 0, self.empty(0, torch.int64), [self.empty(0, torch.int64) for _ in range(na)]
        kp = None if keys is None else keys.data_ptr()
        while True:
            ok = self.empty(cap, torch.int64); oa = self.empty(max(na, 1) * cap, torch.int64)
            specs = (AggSpec * max(na, 1))(*[AggSpec(c.data_ptr() if c is not None and n else None, f, t) for c, f, t in zip(cols, funcs, types)])
            g = C.c_uint64()
            st = self.lib.olap_group_by_i64(self.ctx.handle, kp, specs, na, n, hint,
                                            None, 0, ok.data_ptr(), oa.data_ptr(), cap, None, C.byref(g), self._stream())
            if st == _lib.ERR_CAPACITY:
                cap = int(g.value); continue
            _lib.check(st)
            return g.value, ok[:g.value], [oa[i * cap:i * cap + g.value] for i in range(na)]


class _Growable:
    """Lazily sized view of a cached buffer: out_bufs[i][:total] allocates / reuses capacity once total is known."""

    def __init__(self, buf, name, dtype):
        self.buf, self.name, self.dtype = buf, name, dtype

    def __getitem__(self, sl):
        return self.buf(self.name, sl.stop, self.dtype)


class ShardedExchange:
    """all-to-all of variable-size per-destination segments (the only collective on the path)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.bits = _log2(self.world)
        self.bytes_sent = 0
        self.nccl = dist.is_initialized() and dist.get_backend(group) == "nccl"

    def exchange_counts(self, send_counts):
        if self.world == 1:
            return list(send_counts)
        s = torch.tensor(send_counts, dtype=torch.int64, device="cuda" if self.nccl else "cpu")
        r = torch.empty_like(s)
        dist.all_to_all_single(r, s, group=self.group)
        return r.tolist()

    def all_to_all(self, tensor, send_counts, recv_counts, out=None):
        """Contiguous variable-size all-to-all (small payloads: partial aggregates)."""
        if self.world == 1:
            return tensor
        if out is None:
            out = torch.empty(sum(recv_counts), dtype=tensor.dtype, device=tensor.device)
        dist.all_to_all_single(out, tensor, output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts), group=self.group)
        self.bytes_sent += (sum(send_counts) - send_counts[self.rank]) * tensor.element_size()
        return out

    def exchange_regions(self, cols, begin, count, out_bufs=None, skip_self=False, prefix=None):
        """Ship region p = cols[i][begin[p] : begin[p] + count[p]] to rank p, for every column in cols.

        The bulk goes as multiple-of-4-row segments from 16-row-aligned starts into 4-row-aligned receive offsets: both
        ends of every NCCL peer copy are 16-byte aligned (425 GB/s instead of ~170 GB/s measured on B200 / NVLink 5).
        The <= 3 leftover rows per peer follow in one tiny all-to-all.  Returns ([received columns], segments) with
        segments = [(source rank, rows)] in buffer order.
        skip_self: this rank's own region stays where it is (the caller reads it in place).  prefix[i]: a tensor copied to
        the front of column i's receive buffer (padded to 4 rows); every column's received rows then start at the same,
        4-row-aligned offset `base`, and the segment list starts with (own rank, base)."""
        W = self.world
        if W == 1:
            return [c[begin[0]:begin[0] + count[0]] for c in cols], [(0, count[0])]
        count = list(count)
        if skip_self:
            count[self.rank] = 0
        recv = self.exchange_counts(count)
        c4 = [c & ~3 for c in count]; r4 = [c & ~3 for c in recv]
        cl = [c - d for c, d in zip(count, c4)]; rl = [c - d for c, d in zip(recv, r4)]
        total = sum(recv)
        npre = max([int(t.numel()) for t in prefix if t is not None], default=0) if prefix else 0
        base = (npre + 3) & ~3
        outs = []
        for i, col in enumerate(cols):
            out = out_bufs[i][:base + total] if out_bufs else torch.empty(base + total, dtype=col.dtype, device=col.device)
            if prefix and prefix[i] is not None:
                out[:prefix[i].numel()].copy_(prefix[i])
            ins = [col[begin[p]:begin[p] + c4[p]] for p in range(W)]
            off, views = base, []
            for sidx in range(W):
                views.append(out[off:off + r4[sidx]]); off += r4[sidx]
            if self.nccl:
                dist.all_to_all(views, ins, group=self.group)
            else:  # gloo has no list all_to_all: same data movement through the contiguous form
                tmp = torch.empty(sum(r4), dtype=col.dtype, device=col.device)
                dist.all_to_all_single(tmp, torch.cat(ins), output_split_sizes=r4, input_split_sizes=c4, group=self.group)
                out[base:base + sum(r4)].copy_(tmp)
            left = torch.cat([col[begin[p] + c4[p]:begin[p] + count[p]] for p in range(W)])
            dist.all_to_all_single(out[off:base + total], left, output_split_sizes=rl, input_split_sizes=cl, group=self.group)
            self.bytes_sent += (sum(count) - count[self.rank]) * col.element_size()
            outs.append(out)
        segs = [(sidx, r4[sidx]) for sidx in range(W)] + [(sidx, rl[sidx]) for sidx in range(W)]
        if base:
            segs = [(self.rank, base)] + segs
        return outs, segs


class PeerExchange:
    """The shuffle's data movement over NVSwitch peer memory instead of NCCL point-to-point: every rank owns a symmetric
    receive buffer per column (torch symmetric memory: the peers' buffers are mapped into this process), the W x W count
    matrix is all-gathered (tiny), and each rank then stores its region for rank d directly at its final offset in d's
    buffer -- a dense, contiguous receive buffer, no alignment rules, 766 GB/s per direction measured on B200 / NVLink 5
    against 425 GB/s for NCCL send/recv (scratch/symm_probe.py).  A device-side barrier on the stream orders the stores
    before the readers and the readers of the previous step before the stores."""

    def __init__(self, group=None):
        import torch.distributed._symmetric_memory as symm
        self.symm = symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.bufs = {}          # name -> (local tensor, handle, capacity rows)
        self.bytes_sent = 0

    def _buffer(self, name, rows, dtype, device):
        ent = self.bufs.get(name)
        if ent is None or ent[2] < rows:
            cap = int(rows)
            t = self.symm.empty(cap, dtype=dtype, device=device)
            h = self.symm.rendezvous(t, self.group)
            ent = (t, h, cap)
            self.bufs[name] = ent
        return ent

    def exchange(self, cols, names, begin, count, capacity_rows):
        """cols[i][begin[d] : begin[d] + count[d]] goes to rank d.  capacity_rows must be the same on every rank; returns
        None (on every rank) when some rank would receive more than that -- the caller then takes the NCCL path."""
        W, r = self.world, self.rank
        dev = cols[0].device
        mine = torch.tensor(count, dtype=torch.int64, device=dev)
        allc = torch.empty(W * W, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, mine, group=self.group)
        Cm = allc.view(W, W).tolist()                       # Cm[s][d]: rows s sends to d (host waits: tiny)
        recv_tot = [sum(Cm[s][d] for s in range(W)) for d in range(W)]
        if max(recv_tot) > capacity_rows:
            return None
        outs = []
        for col, name in zip(cols, names):
            t, h, cap = self._buffer(name, capacity_rows, col.dtype, dev)
            h.barrier()                                     # everyone is done reading what the previous step delivered
            for k in range(W):
                d = (r + k) % W                             # start with the own region, then round-robin over the peers
                n = Cm[r][d]
                if n == 0:
                    continue
                off = sum(Cm[s2][d] for s2 in range(r))
                dst = t if d == r else h.get_buffer(d, (cap,), col.dtype)
                dst[off:off + n].copy_(col[begin[d]:begin[d] + n])
                if d != r:
                    self.bytes_sent += n * col.element_size()
            h.barrier()                                     # all stores have landed
            outs.append(t[:recv_tot[r]])
        return outs, [(s2, Cm[s2][r]) for s2 in range(W)]


    def fused_shuffle(self, ops, keys, n, names, capacity_rows):
        """Partition + exchange in one kernel: the partition kernel of every rank stores (key, source row id) of each row
        straight into the destination rank's receive buffers at the final offset (olap_shuffle_scatter_i64).  Only the
        W x W count matrix goes through a collective.  Returns like exchange(); None when a rank would overflow."""
        W, r = self.world, self.rank
        bits = _log2(W)
        dev = keys.device
        lib, ctx = ops.lib, ops.ctx
        words = int(lib.olap_shuffle_scratch_words(ctx.handle, bits))
        scratch = ops.buffer("shuffle_scratch_" + names[0], words, torch.int32)
        cnt = torch.zeros(W, dtype=torch.int32, device=dev)
        if n:   # a rank without rows still takes part in every collective below
            _lib.check(lib.olap_shuffle_count_i64(ctx.handle, keys.data_ptr(), n, bits, scratch.data_ptr(), cnt.data_ptr(), ops._stream()))
        allc = torch.empty(W * W, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, cnt.to(torch.int64), group=self.group)
        Cm = allc.view(W, W).tolist()                       # Cm[s][d]: rows s sends to d (host waits: tiny)
        recv_tot = [sum(Cm[s][d] for s in range(W)) for d in range(W)]
        if max(recv_tot) > capacity_rows:
            return None
        tk, hk, _ = self._buffer(names[0], capacity_rows, torch.int64, dev)
        ti, hi, _ = self._buffer(names[1], capacity_rows, torch.int32, dev)
        hk.barrier()                                        # everyone is done reading what the previous step delivered
        dk = (C.c_void_p * W)(*[int(p) for p in hk.buffer_ptrs])
        di = (C.c_void_p * W)(*[int(p) for p in hi.buffer_ptrs])
        off = (C.c_uint64 * W)(*[sum(Cm[s2][d] for s2 in range(r)) for d in range(W)])
        if n:
            _lib.check(lib.olap_shuffle_scatter_i64(ctx.handle, keys.data_ptr(), n, bits, scratch.data_ptr(), dk, di, off, ops._stream()))
        hk.barrier()                                        # every rank's stores have landed
        self.bytes_sent += (n - Cm[r][r]) * 12
        return [tk[:recv_tot[r]], ti[:recv_tot[r]]], [(s2, Cm[s2][r]) for s2 in range(W)]


class ShardedJoin:
    """GpuHashJoin across ranks.  join() returns a dict describing this rank's shard of the result."""

    def __init__(self, ops, group=None):
        self.ops = ops
        self.x = ShardedExchange(group)
        self._comm = None   # side stream for the exchange, so it overlaps the next partition kernel
        self.fuse = os.environ.get("OLAP_FUSED_SHUFFLE", "1") != "0"   # partition kernel stores straight into the peers' buffers
        self.peer = None    # NVSwitch peer-memory exchange (CUDA + NCCL group only); OLAP_PEER_EXCHANGE=0 keeps NCCL send/recv
        if self.x.nccl and self.x.world > 1 and os.environ.get("OLAP_PEER_EXCHANGE", "1") != "0":
            try:
                self.peer = PeerExchange(group)
            except Exception as e:  # symmetric memory not available on this build / topology
                print(f"[sharded] peer-memory exchange unavailable ({e}); using NCCL send/recv", file=sys.stderr)

    def _exchange(self, pk, pid, begin, count, tag, capacity_rows=None):
        if self.peer is not None and capacity_rows:
            got = self.peer.exchange([pk, pid], ["recv_k" + tag, "recv_i" + tag], begin, count, capacity_rows)
            if got is not None:
                (rk, ri), segs = got
                self.x.bytes_sent += self.peer.bytes_sent; self.peer.bytes_sent = 0
                return rk, ri, segs
        buf = getattr(self.ops, "buffer", None)
        bufs = None
        if buf and self.x.world > 1:   # cached receive buffers, sized once the counts are known
            bufs = [_Growable(buf, "recv_k" + tag, pk.dtype), _Growable(buf, "recv_i" + tag, pid.dtype)]
        (rk, ri), segs = self.x.exchange_regions([pk, pid], begin, count, bufs)
        return rk, ri, segs

    @staticmethod
    def _capacity(n_local):
        """Rows a rank may receive for one side on the peer path (same formula on every rank): 30 % above an even split."""
        return int(n_local * 1.3) + (1 << 16)

    def _fused(self, keys, n, tag):
        """Fused partition + peer-store shuffle of one side, or None when it is unavailable / would overflow."""
        if self.peer is None or not self.fuse or self.x.world > 8 or not hasattr(self.ops, "lib"):
            return None
        # the capacity (it sizes a symmetric buffer: a collective) and the go / no-go decision must be the same on every rank:
        # both derive from the LARGEST shard, not from this rank's row count
        nmax = torch.tensor([n], dtype=torch.int64, device=keys.device)
        dist.all_reduce(nmax, op=dist.ReduceOp.MAX, group=self.x.group)
        nmax = int(nmax.item())
        if nmax == 0:
            return None
        got = self.peer.fused_shuffle(self.ops, keys, n, ["recv_k" + tag, "recv_i" + tag], self._capacity(nmax))
        if got is None:
            return None
        (rk, ri), segs = got
        self.x.bytes_sent += self.peer.bytes_sent; self.peer.bytes_sent = 0
        return rk, ri, segs

    def shuffle(self, keys, n, tag=""):
        got = self._fused(keys, n, tag)
        if got is not None:
            return got
        pk, pid, begin, count = self.ops.shuffle_partition(keys, n, self.x.bits, tag)
        return self._exchange(pk, pid, begin, count, tag, None)   # NCCL path: no symmetric buffer, so no per-rank capacity to agree on

    def shuffle_both(self, build_keys, n_build, probe_keys, n_probe):
        """Both sides; on CUDA the build-side exchange runs on a side stream while the probe side is being partitioned."""
        if not (build_keys.is_cuda and self.x.world > 1 and hasattr(self.ops, "shuffle_partition_async")) or (self.peer is not None and self.fuse):
            return self.shuffle(build_keys, n_build, "_b"), self.shuffle(probe_keys, n_probe, "_p")
        P = self.x.world
        main = torch.cuda.current_stream()
        if self._comm is None:
            self._comm = torch.cuda.Stream()
        bk_s, bi_s, bbegin, bcount = self.ops.shuffle_partition(build_keys, n_build, self.x.bits, "_b")   # host waits for the build partition
        done_b = torch.cuda.Event(); done_b.record(main)
        pk_s, pi_s, pmeta = self.ops.shuffle_partition_async(probe_keys, n_probe, self.x.bits, "_p")      # queued, not awaited
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(done_b)
            b = self._exchange(bk_s, bi_s, bbegin, bcount, "_b", self._capacity(n_build))
            ev = torch.cuda.Event(); ev.record(self._comm)
        m = pmeta.tolist()                              # host waits for the probe partition; the build exchange ran meanwhile
        main.wait_event(ev)
        p = self._exchange(pk_s, pi_s, m[:P + 1], m[P + 1:], "_p", self._capacity(n_probe))
        return b, p

    def join(self, build_keys, n_build, probe_keys, n_probe, max_matches=None, strategy=JOIN_AUTO):
        if build_keys.is_cuda and self.x.world > 1 and hasattr(self.ops, "join_build") and self.peer is None:
            return self._join_overlapped(build_keys, n_build, probe_keys, n_probe, max_matches, strategy)
        (bk, bi, bsegs), (pk, pi, psegs) = self.shuffle_both(build_keys, n_build, probe_keys, n_probe)
        nb, npb = int(bk.numel()), int(pk.numel())
        mm = max_matches if max_matches is not None else max(npb, 1)
        count, pairs = self.ops.hash_join(bk, nb, pk, npb, mm, strategy)
        return {"count": count, "pairs": pairs, "build_ids": bi, "build_recv": bsegs, "probe_ids": pi, "probe_recv": psegs,
                "n_build_local": nb, "n_probe_local": npb}

    def _join_overlapped(self, build_keys, n_build, probe_keys, n_probe, max_matches, strategy):
        """The probe rows this rank keeps are joined in place (first probe batch) while the rows of the other ranks are
        still on the wire (second batch); the build side is prepared once.  The virtual probe receive buffer the pairs
        index is [own rows, padded to 4 | received rows]."""
        P, r = self.x.world, self.x.rank
        main = torch.cuda.current_stream()
        if self._comm is None:
            self._comm = torch.cuda.Stream()
        comm = self._comm
        dbg = os.environ.get("OLAP_SHARD_DEBUG")
        marks = []

        def mark(name, stream):
            if dbg:
                e = torch.cuda.Event(enable_timing=True); e.record(stream); marks.append((name, e))

        mark("start", main)
        bk_s, bi_s, bbegin, bcount = self.ops.shuffle_partition(build_keys, n_build, self.x.bits, "_b")   # host waits for the build partition
        done_b = torch.cuda.Event(); done_b.record(main)
        mark("build partitioned", main)
        pk_s, pi_s, pmeta = self.ops.shuffle_partition_async(probe_keys, n_probe, self.x.bits, "_p")      # queued, not awaited
        mark("probe partitioned", main)
        with torch.cuda.stream(comm):
            comm.wait_event(done_b)
            rk_b, ri_b, bsegs = self._exchange(bk_s, bi_s, bbegin, bcount, "_b")
            ev_b = torch.cuda.Event(); ev_b.record(comm)
            mark("build exchanged (comm)", comm)
        m = pmeta.tolist()                              # host waits for the probe partition; the build exchange ran meanwhile
        pbegin, pcount = m[:P + 1], m[P + 1:]
        n_own = pcount[r]
        own_ids = pi_s[pbegin[r]:pbegin[r] + n_own]
        with torch.cuda.stream(comm):                   # the other ranks' probe rows travel while this rank builds and probes its own
            bufs = [_Growable(self.ops.buffer, "recv_k_p", pk_s.dtype), _Growable(self.ops.buffer, "recv_i_p", pi_s.dtype)]
            (rk, ri), psegs = self.x.exchange_regions([pk_s, pi_s], pbegin, pcount, bufs, skip_self=True, prefix=[None, own_ids])
            ev_p = torch.cuda.Event(); ev_p.record(comm)
            mark("probe exchanged (comm)", comm)
        base = (n_own + 3) & ~3
        n_remote = int(rk.numel()) - base
        main.wait_event(ev_b)
        nb_loc = int(rk_b.numel())
        mm = max_matches if max_matches is not None else max(n_own + n_remote, 1)
        pairs = self.ops.buffer("pairs", 2 * max(mm, 1), torch.int32)
        jb = self.ops.join_build(rk_b, nb_loc, strategy)
        mark("build side prepared", main)
        try:
            c0 = self.ops.join_probe(jb, pk_s.data_ptr() + 8 * pbegin[r], n_own, 0, pairs.data_ptr(), mm)
            mark("own rows joined", main)
            main.wait_event(ev_p)
            c1 = self.ops.join_probe(jb, rk.data_ptr() + 8 * base, n_remote, base, pairs.data_ptr() + 8 * c0, mm - c0)
            mark("received rows joined", main)
        finally:
            jb.close()
        if dbg and r == 0:
            torch.cuda.synchronize()
            print("[sharded join] " + ", ".join(f"{n} @{marks[0][1].elapsed_time(e):.2f} ms" for n, e in marks[1:]), file=sys.stderr, flush=True)
        return {"count": c0 + c1, "pairs": pairs, "build_ids": ri_b, "build_recv": bsegs, "probe_ids": ri, "probe_recv": psegs,
                "n_build_local": nb_loc, "n_probe_local": n_own + n_remote}

    @staticmethod
    def global_pairs(res, build_rank_offsets, probe_rank_offsets):
        """Translate (receive-buffer positions) into global (left_row, right_row) ids; int64 tensor [count, 2]."""
        k = int(res["count"])
        pairs = res["pairs"][:2 * k].view(-1, 2).to(torch.int64) & 0xFFFFFFFF

        def to_global(pos, ids, segs, offs):
            bounds = torch.tensor(np.cumsum([0] + [c for _, c in segs]), device=pos.device)
            seg = torch.bucketize(pos.contiguous(), bounds[1:], right=True)
            seg_off = torch.tensor([offs[src] for src, _ in segs], device=pos.device, dtype=torch.int64)
            return seg_off[seg] + (ids.to(torch.int64) & 0xFFFFFFFF)[pos]

        left = to_global(pairs[:, 0], res["build_ids"], res["build_recv"], build_rank_offsets)
        right = to_global(pairs[:, 1], res["probe_ids"], res["probe_recv"], probe_rank_offsets)
        return torch.stack([left, right], 1)


class ShardedGroupBy:
    """GpuAggregate across ranks: local partial aggregates, key-sharded exchange of the partials, per-shard merge."""

    def __init__(self, ops, group=None):
        self.ops = ops
        self.x = ShardedExchange(group)

    def group_by(self, keys, n, cols, funcs, types, n_groups_hint):
        # 1. local pre-aggregation.  Partials are deduplicated: AVG(x) and SUM(x) share one sum, every COUNT/AVG shares the
        #    one row count -- so the merge needs at most as many accumulators as the local pass
        pf, pc, pt, index = [], [], [], {}

        def partial(f, c, t):
            key = ("count",) if f == COUNT else (f, id(c), t)
            if key not in index:
                index[key] = len(pf)
                pf.append(f); pc.append(c); pt.append(I64 if f == COUNT else t)
            return index[key]

        plan = []
        for c, f, t in zip(cols, funcs, types):
            plan.append((partial(SUM, c, t), partial(COUNT, c, t)) if f == AVG else (partial(f, c, t), None))
        g, gk, parts = self.ops.group_by(keys, n, pc, pf, pt, n_groups_hint)
        # 2. route every partial row to the rank that owns its key
        pk, pid, begin, send = self.ops.shuffle_partition(gk, g, self.x.bits, "_g")
        recv = self.x.exchange_counts(send)
        # partials are tiny: compact the (16-row aligned) regions and use the plain contiguous all-to-all
        sel = torch.cat([torch.arange(begin[p], begin[p] + send[p], device=pk.device) for p in range(self.x.world)])
        order = pid.to(torch.int64)[sel] & 0xFFFFFFFF
        rk = self.x.all_to_all(pk[sel].contiguous(), send, recv)
        rparts = [self.x.all_to_all(p.index_select(0, order).contiguous(), send, recv) for p in parts]
        # 3. merge: SUM/COUNT add, MIN/MAX combine
        mf = [SUM if f in (SUM, COUNT) else f for f in pf]
        m, mk, merged = self.ops.group_by(rk, int(rk.numel()), rparts, mf, pt, max(n_groups_hint, 1))
        # 4. finalise
        out = []
        for (f, t), (i, ic) in zip(zip(funcs, types), plan):
            if f == AVG:
                s = merged[i].view(torch.float64) if t == F64 else merged[i].to(torch.float64)
                cnt = merged[ic].to(torch.float64)
                out.append(torch.where(cnt > 0, s / cnt, torch.zeros_like(s)))
            else:
                out.append(merged[i].view(torch.float64) if (t == F64 and f != COUNT) else merged[i])
        return m, mk, out


# ------------------------------------------------------------------------------------------------- bench (N > 1)
def _slab_views(rj, torch):
    """(build words, probe words) of this rank's receive slab as int64 tensors, plus the layout."""
    lay = (C.c_uint64 * 8)()
    _lib.check(rj.lib.olap_shard_layout(rj.handle, lay))
    words_b, words_p = int(lay[0]), int(lay[1])
    w = rj.slab.view(torch.int64)
    return w[:words_b], w[words_b:words_b + words_p * rj.chunks], [int(x) for x in lay]


def verify_range_join(rj, torch, dist, pairs, count, npb_global, hot_count=0):
    """Every pair of this rank, on the device: the build word and the probe word the pair points at carry the same key offset,
    no probe position appears twice, and (all probe keys match exactly one build key) the global count is the global probe rows."""
    bw, pw, lay = _slab_views(rj, torch)
    pr = pairs[:2 * count].view(-1, 2)
    seen = torch.zeros(pw.numel(), dtype=torch.uint8, device=pairs.device)
    step = 1 << 26
    for a in range(0, count, step):
        l = pr[a:a + step, 0].to(torch.int64) & 0xFFFFFFFF
        r = pr[a:a + step, 1].to(torch.int64) & 0xFFFFFFFF
        assert bool(((bw[l] >> 32) == (pw[r] >> 32)).all()), "sharded join: pair with different keys"
        seen[r] = 1
    assert int(seen.sum(dtype=torch.int64)) == count, "sharded join: a received probe row appears twice"
    tot = torch.tensor([count + hot_count], dtype=torch.int64, device=pairs.device)
    dist.all_reduce(tot)
    assert int(tot.item()) == npb_global, f"global join count {int(tot.item())} != {npb_global}"


def bench_sharded_join(ctx, args, rank, world, local_rank, peak, peak_kind, ClockSampler, cpu_baseline=None):
    """bench.py's N > 1 leg: weak scaling, per-GPU 2^log2_build unique-key build x 2^log2_probe FK probe over the GLOBAL key
    domain; rows are contiguous-sharded (not pre-partitioned by key).  value = rows of all ranks / max-over-ranks time."""
    from .shard import RangeShardedJoin
    ops = DeviceOps(ctx)
    dev = ops.device
    nb, npb = 1 << args.log2_build, 1 << args.log2_probe
    gnb = nb * world
    chunks = int(os.environ.get("OLAP_SHARD_CHUNKS", "8"))
    build = ops.empty(nb, torch.int64); probe = ops.empty(npb, torch.int64)
    ctx.generate(nb, 1, 3, gnb, row0=rank * nb, out=build.data_ptr())        # slice of a permutation of [0, world * 2^log2_build)
    zipf = float(getattr(args, "zipf", 0.0) or 0.0)
    if zipf > 0:
        ctx.generate_fk(npb, 3, gnb, 33, zipf, row0=rank * npb, out=probe.data_ptr())   # FK ~ Zipf(s) over the global build rows
    else:
        ctx.generate(npb, 0, 33, gnb, row0=rank * npb, out=probe.data_ptr())            # uniform FK over the global build domain
    ctx.synchronize()
    rj = RangeShardedJoin(ctx, nb, npb, chunks)
    mm = int(npb * (2.5 if zipf > 0 else 1.25)) + 4096       # skew: the owners of warm (not hot) keys receive more than an even share
    pairs = torch.empty(2 * mm, dtype=torch.int32, device=dev)
    hot_pairs = torch.empty(2 * npb, dtype=torch.int32, device=dev) if zipf > 0 else None

    def step():
        res = rj.join(build, nb, probe, npb, pairs, mm, hot=1 if zipf > 0 else 0, hot_pairs=hot_pairs, max_hot_pairs=npb)
        assert res["supported"] and not any(res["flags"]), f"range-sharded join refused / flagged: {res}"
        return res

    for _ in range(args.warmup):
        res = step()
    verify_range_join(rj, torch, dist, pairs, res["count"], npb * world, res["hot_count"])     # all pairs of every rank, after warm-up
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count()
    dist.barrier(); torch.cuda.synchronize()
    ev0.record(rj.comm)
    for _ in range(args.steps):
        res = step()
    rj.comm.wait_stream(rj.comp)
    ev1.record(rj.comm)
    torch.cuda.synchronize(); dist.barrier()
    launches = ctx.launch_count() - launches0
    phases = {k: v for k, v in ctx.last_phase_times().items() if k != "end"}
    verify_range_join(rj, torch, dist, pairs, res["count"], npb * world, res["hot_count"])     # and of the last timed step
    sent = (C.c_uint64 * world)()
    _lib.check(rj.lib.olap_shard_sent_rows(rj.handle, sent, rj.comp.cuda_stream))
    bytes_out = 8 * sum(int(sent[d]) for d in range(world) if d != rank)

    # the exchange alone (both sides scattered into the peers' slabs, no local work), >= 5 warmed repetitions, for the NVLink figure
    def shuffle_only():
        lib, h, comm = rj.lib, rj.handle, rj.comm
        gmin, gmax, gn = rj._global_range(build, nb)
        _lib.check(lib.olap_shard_plan(h, gmin, gmax, gn, 0, comm.cuda_stream))
        rj._barrier()
        if rj.part is not comm:
            ev = torch.cuda.Event(); ev.record(comm); rj.part.wait_event(ev)
        rj._scatter(0, 0, build.data_ptr(), nb, 0)
        cc = rj.chunk_cap
        for k in range(rj.chunks):
            lo = min(k * cc, npb); hi = min(lo + cc, npb)
            rj._scatter(1, k, probe.data_ptr() + 8 * lo, hi - lo, lo)
        rj._barrier()

    # ... and the local work alone (second partition level + shared-memory join of every chunk, on what the last shuffle delivered)
    def local_only():
        lib, h, comp = rj.lib, rj.handle, rj.comp
        _lib.check(lib.olap_shard_build(h, comp.cuda_stream))
        for k in range(rj.chunks):
            _lib.check(lib.olap_shard_probe(h, k, pairs.data_ptr(), mm, comp.cuda_stream))

    torch.cuda.synchronize()
    for _ in range(2):
        shuffle_only()
    torch.cuda.synchronize()
    local_only()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    l0.record(rj.comp)
    for _ in range(3):
        local_only()
    l1.record(rj.comp)
    torch.cuda.synchronize()
    local_ms = l0.elapsed_time(l1) / 3
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record(rj.comm)
    for _ in range(reps):
        shuffle_only()
    e1.record(rj.comm)
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps, e0.elapsed_time(e1) / reps, local_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step, shuffle_ms, local_ms = ms[0].item(), ms[1].item(), ms[2].item()
    rows = (nb + npb) * world

    # e2e: host (pinned) key columns -> H2D -> sharded join -> D2H of this rank's pairs; wall clock, max over ranks
    e2e_note = None
    try:
        if args.skip_e2e:
            raise RuntimeError("skipped (--skip-e2e)")
        hb = torch.empty(nb, dtype=torch.int64, pin_memory=True); hp = torch.empty(npb, dtype=torch.int64, pin_memory=True)
        hb.copy_(build); hp.copy_(probe)
        hout = torch.empty(2 * (res["count"] + 1024), dtype=torch.int32, pin_memory=True)
        db2, dp2 = torch.empty_like(build), torch.empty_like(probe)
    except Exception as e:  # e.g. the host cannot pin 20 GB per rank: the line then says so instead of dying
        e2e_note = f"no end-to-end leg: {e}"
        args.skip_e2e = True
    flag = torch.tensor([1 if args.skip_e2e else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)       # one rank without host buffers = no rank runs the leg
    args.skip_e2e = bool(flag.item())

    def e2e_step():
        db2.copy_(hb, non_blocking=True); dp2.copy_(hp, non_blocking=True)
        r = rj.join(db2, nb, dp2, npb, pairs, mm)
        hout[:2 * r["count"]].copy_(pairs[:2 * r["count"]], non_blocking=True)
        torch.cuda.synchronize()
        return r["count"]

    n_e2e = max(1, min(args.steps, 2))
    out_rows = 0
    t0 = time.perf_counter()
    if not args.skip_e2e:
        e2e_step()
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            out_rows = e2e_step()
        dist.barrier()
    e2e = torch.tensor([(time.perf_counter() - t0) * 1e3 / n_e2e], dtype=torch.float64, device=dev)
    dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        lay = (C.c_uint64 * 8)()
        _lib.check(rj.lib.olap_shard_layout(rj.handle, lay))
        nvl = bytes_out / (shuffle_ms * 1e-3) / 1e9 if shuffle_ms else None
        cpu = cpu_baseline() if cpu_baseline and not args.skip_cpu else None
        line = {
            "metric": "join rows/sec (|build|+|probe|)/time", "value": rows / (ms_step * 1e-3), "unit": "rows/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"exchange_mode": "ce" if rj.ce_mode else "tma",
                       "workload": f"{world}-GPU key-range-sharded radix hash join, per GPU 2^{args.log2_build} unique-key build x 2^{args.log2_probe} FK probe "
                                   f"(global 2^{args.log2_build}*{world} x 2^{args.log2_probe}*{world}), int64 keys, " + (f"Zipf s = {zipf}" if zipf > 0 else "uniform (Zipf s = 0)") + ", rows contiguous-sharded",
                       "exchange": (f"dense key domain: one contiguous key slice per rank ({int(lay[4])} sub-ranges of {int(lay[5])} x 2^15 keys); rows travel as packed 8-byte words "
                                    f"{{key - slice start, source row id}} into fixed-capacity single-writer regions of the owner's receive slab (torch symmetric memory, peer-mapped), no count pass; "
                                    + ("copy-engine mode: the partition kernel fills a local outbox laid out like the receivers' regions (own rows go straight into the own slab), one peer "
                                       "cudaMemcpyAsync per destination moves each block -- NVLink runs next to the SMs' partition + join work; "
                                       if rj.ce_mode else
                                       "TMA mode: ONE kernel partitions and ships -- every run leaves the shared-memory stage with cp.async.bulk into the peer's slab, one CTA per SM next to the join kernels; ")
                                    + f"probe side in {chunks} chunks, chunk k+1 on the wire while chunk k is partitioned and joined in shared memory; barriers are a one-warp kernel signalling "
                                    "through the slabs; NCCL carries only the min/max/flag reductions"),
                       "l2": "inputs larger than L2", "parallelism": f"key-range shards x{world}",
                       "verified": "all pairs of every rank after warm-up and after the timed steps (equal keys, each received probe row once, global count)"},
            "clocks": clocks,
            "e2e": {"value": rows / (e2e.item() * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": (nb + npb) * 8 * world,
                    "d2h_bytes_per_step": int(out_rows) * 8 * world, "ms_per_step": e2e.item(), "steps": n_e2e,
                    "how": e2e_note or "pinned host key columns -> H2D -> range-sharded join -> D2H of every rank's pairs; wall clock, max over ranks"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "nvlink", "kernel": "rpart_kernel<raw, peer> (shuffle: both sides of one join scattered into the peers' slabs, measured alone)",
                         "achieved": nvl, "peak": 770.0, "peak_kind": "measured peer copy per direction (B200_PROFILING.md); nominal 900", "unit": "GB/s",
                         "frac": nvl / 770.0 if nvl else None, "traffic": None,
                         "algorithmic_bytes_per_launch": bytes_out, "ms_per_launch": shuffle_ms, "repetitions": reps},
            "nvlink": {"shuffle_ms_per_join": shuffle_ms, "bytes_out_per_rank_per_join": bytes_out, "achieved_gbs_per_direction": nvl,
                       "peak_nominal_gbs": 900.0, "peak_measured_gbs": 770.0,
                       "floor_ms_per_join_at_measured_peak": bytes_out / 770e9 * 1e3,
                       "note": "8 bytes leave the rank per row whose key another rank owns; inside the step the shuffle of chunk k+1 overlaps the local join of chunk k"},
            "skew_note": (None if zipf <= 0 else
                          "Zipf: the timed steps run WITH heavy-hitter handling (hot keys agreed across ranks and remembered per probe column, their build rows replicated, "
                          "hot probe rows joined on their source rank inside the shuffle pass); the shuffle-alone / local-alone figures below are taken with it switched off "
                          "(every row travels), so `hidden_ms` and the NVLink rate are not comparable with the step -- compare ms_per_step with the uniform run"),
            "overlap": {"shuffle_alone_ms": shuffle_ms, "local_alone_ms": local_ms, "step_ms": ms_step,
                        "hidden_ms": shuffle_ms + local_ms - ms_step,
                        "note": "shuffle = both sides scattered; local = second partition level + shared-memory join of all chunks; hidden = how much of their sum the step does not pay"},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    dist.barrier()
    rj.close()
    dist.destroy_process_group()
    return 0
