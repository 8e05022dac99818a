#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 execution core (contract: see the task prompt / DESIGN.md "Measurement").

Metric (BASELINE.json): join rows/sec, rows = |build| + |probe|, on the single-GPU radix hash join
2^27 unique-key build x 2^29 FK probe, int64 keys (configs[1]); % of HBM roofline for the dominant kernel.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2-build B --log2-probe P] [--zipf S]

* value     device-resident inputs, CUDA events around K joins on the launching stream, max over ranks; every pair verified afterwards
* e2e       the same join through the C-ABI with HOST (pinned) buffers: H2D of both key columns, join, D2H of the
            count and of the row-id pairs, all inside the timed region
* roofline  the kernel with the largest share of the step (djoin_kernel on the dense-key path): algorithmic bytes / its CUDA-event
            duration in the timed region, against MEASURED_PEAKS.json's hbm_gbs; roofline_kernels lists every phase the same way
* value_with_payload / value_general / group_by   the same line also carries the payload-materialising join (BASELINE config 2's "int64
            payload"), the sparse-key (bucket-table) path, and BASELINE config 3's group-by sweep (G = 1e2 / 1e5 / 1e8 at 2^30 rows)
* cpu_baseline / --impl reference   the oracle's multi-threaded port of the reference CPU path on this box's cores
N > 1: weak scaling, one process per GPU (torchrun), per GPU 2^28 x 2^30 (BASELINE config 5); the join is key-RANGE-sharded over NVLink
peer memory (gpu_olap_engine_b200/shard.py + sharded.py::bench_sharded_join), NCCL carries only the tiny reductions;
value = rows of all ranks / max-over-ranks time; --zipf S skews the probe keys (heavy hitters are replicated).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only when MEASURED_PEAKS.json is absent


def peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill(); out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load": samples whose power is in the upper half of what was seen during the timed region
        thr = (max(pw) + min(pw)) / 2
        load = [s for s, p in zip(sm, pw) if p >= thr] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def cpu_reference_join(log2_build, log2_probe, threads, steps=1, warmup=0):
    """The reference's CPU path (oracle port: executor.rs:238-253 build + the omitted probe), radix-partitioned over
    all host threads.  Returns (rows_per_sec, seconds_per_step, count)."""
    import numpy as np
    import oracle as O
    nb, npb = 1 << log2_build, 1 << log2_probe
    b = O.gen_perm_i64(nb, 3, nb); p = O.gen_uniform_i64(npb, 33, nb)
    times = []
    cnt = 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        cnt, pairs = O.hash_join_i64(b, p, max_matches=npb, threads=threads)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    assert cnt == npb
    t = sum(times) / len(times)
    return (nb + npb) / t, t, cnt


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    if args.gpus > 1 and not args.sizes_given:
        args.log2_build, args.log2_probe = 28, 30   # the arm it is compared with (BASELINE config 5 per GPU)
    lb, lp = args.cpu_log2_build, args.cpu_log2_probe
    rps, sec, _ = cpu_reference_join(lb, lp, cores, steps=args.steps, warmup=min(args.warmup, 1))
    sample = (f"bounded sample of the workload: 2^{lb} unique-key build x 2^{lp} FK probe, int64 keys, pairs materialised; "
              f"oracle/oracle.c orc_hash_join_i64_mt (C port of executor.rs:238-253 + probe; the Rust reference cannot be built here)")
    line = {
        "impl": "reference", "metric": "join rows/sec (|build|+|probe|)/time", "value": rps, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": (f"single-GPU radix hash join 2^{args.log2_build} unique-key build x 2^{args.log2_probe} FK probe, int64 keys, uniform"
                                if args.gpus <= 1 else
                                f"{args.gpus}-GPU key-sharded radix hash join, per GPU 2^{args.log2_build} unique-key build x 2^{args.log2_probe} FK probe "
                                f"(global 2^{args.log2_build}*{args.gpus} x 2^{args.log2_probe}*{args.gpus}), int64 keys, uniform, rows contiguous-sharded"),
                   "sample": sample},
        "cpu_baseline": {"value": rps, "unit": "rows/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rps, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import gpu_olap_engine_b200 as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1 (one process per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = g.Context(local_rank, 0, 8 if world == 1 else 12)   # N > 1: streams 3.. carry the peer copies, one per destination
    ctx.set_profiling(True)
    peak, peak_kind = peak_hbm()

    if world > 1:
        from gpu_olap_engine_b200 import sharded
        if not args.sizes_given:   # BASELINE config 5: per GPU 2^28 x 2^30 (2^31 x 2^33 at 8 GPUs)
            args.log2_build, args.log2_probe = 28, 30

        def cpu():
            cores = os.cpu_count() or 1
            rps, sec, _ = cpu_reference_join(args.cpu_log2_build, args.cpu_log2_probe, cores)
            return {"value": rps, "unit": "rows/s", "cores": cores, "kind": "port",
                    "sample": f"2^{args.cpu_log2_build} unique-key build x 2^{args.cpu_log2_probe} FK probe (same generator), pairs materialised, {cores} threads, {sec:.2f} s"}
        return sharded.bench_sharded_join(ctx, args, rank, world, local_rank, peak, peak_kind, ClockSampler, cpu)

    return bench_single(ctx, g, args, local_rank, peak, peak_kind)


# Algorithmic bytes per row of every phase the join can run: what the kernel must read + write once.
#   smem direct path: rows travel as one packed 8-byte word {key - min, row id}
#   bucket / L2-direct path: 8-byte key + 4-byte id
PHASE_BYTES = {
    "minmax": ("key_minmax_kernel", lambda nb, npb, c: 8.0 * nb),
    "partition_build_coarse": ("rpart_kernel<raw> (build side, 64-range coarse regions)", lambda nb, npb, c: 16.0 * nb),
    "partition_build": ("rpart_kernel<regions> (build side, ranges of 2^15 keys)", lambda nb, npb, c: 16.0 * nb),
    "partition_probe_coarse": ("rpart_kernel<raw> (probe side, 64-range coarse regions)", lambda nb, npb, c: 16.0 * npb),
    "partition_probe": ("rpart_kernel<regions> (probe side, ranges of 2^15 keys)", lambda nb, npb, c: 16.0 * npb),
    "build_probe_smem": ("djoin_kernel (build + probe of every key range in a 128 KB shared-memory table)", lambda nb, npb, c: 8.0 * (nb + npb) + 8.0 * c),
    "build_probe_direct": ("join_pipeline_kernel<direct> (L2-resident row-id tables)", lambda nb, npb, c: 12.0 * (nb + npb) + 8.0 * c),
    "build_probe": ("join_pipeline_kernel<buckets> (L2-resident two-slot bucket tables)", lambda nb, npb, c: 12.0 * (nb + npb) + 8.0 * c),
}
GENERAL_PHASE_BYTES = {
    "partition_build": ("partition_kernel (build side, 256 hash partitions)", lambda nb, npb, c: 20.0 * nb),
    "partition_probe": ("partition_kernel (probe side, 256 hash partitions)", lambda nb, npb, c: 20.0 * npb),
}


def load_traffic():
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def kernel_rooflines(phases, table, nb, npb, count, peak, traffic):
    out = []
    for name, ms in phases.items():
        ent = table.get(name) or PHASE_BYTES.get(name)
        if not ent or not ms:
            continue
        by = ent[1](nb, npb, count)
        gbs = by / (ms * 1e-3) / 1e9
        out.append({"phase": name, "kernel": ent[0], "ms_per_launch": ms, "algorithmic_bytes_per_launch": by, "achieved": gbs, "frac": gbs / peak,
                    "traffic": traffic.get(name)})
    out.sort(key=lambda r: -r["ms_per_launch"])
    return out


def timed_join(ctx, g, torch, stream, dbuild, nb, dprobe, npb, pairs, steps, warmup, expect=None):
    def step():
        c, _ = ctx.hash_join_i64(dbuild, nb, dprobe, npb, max_matches=npb, strategy=g.JOIN_RADIX, out_pairs=pairs)
        return c
    for _ in range(warmup):
        cnt = step()
    assert expect is None or cnt == expect, f"join count {cnt} != {expect}"
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase_acc = {}
    launches0 = ctx.launch_count()
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(steps):
        step()
        for k, v in ctx.last_phase_times().items():
            phase_acc[k] = phase_acc.get(k, 0.0) + v
    ev1.record(stream)
    torch.cuda.synchronize()
    ms_step = ev0.elapsed_time(ev1) / steps
    return ms_step, {k: v / steps for k, v in phase_acc.items() if k != "end"}, (ctx.launch_count() - launches0), cnt


class _DevView:
    """A device buffer of the library as a zero-copy torch tensor (__cuda_array_interface__)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def dev_view(torch, buf, n, typestr="<i8"):
    return torch.as_tensor(_DevView(buf.ptr, n, typestr), device="cuda")


def fk_probe_keys(torch, build_view, out_view, zipf, seed):
    """Probe-side foreign keys: build row r's key, r ~ uniform (zipf = 0) or r + 1 ~ Zipf(s) over the build rows (continuous
    inverse-CDF sampler).  The build keys are a pseudo-random permutation, so hot ranks map to keys spread over the domain
    ("mapped through a fixed permutation")."""
    nb, n = build_view.numel(), out_view.numel()
    gen = torch.Generator(device="cuda"); gen.manual_seed(seed)
    step = 1 << 27
    for a in range(0, n, step):
        m = min(step, n - a)
        if zipf <= 0:
            r = torch.randint(0, nb, (m,), device="cuda", generator=gen)
        else:
            u = torch.rand(m, dtype=torch.float64, device="cuda", generator=gen)
            if abs(zipf - 1.0) < 1e-9:
                r = torch.exp(u * torch.log(torch.tensor(float(nb), dtype=torch.float64, device="cuda")))
            else:
                e = 1.0 - zipf
                r = ((float(nb) ** e - 1.0) * u + 1.0) ** (1.0 / e)
            r = (r.to(torch.int64) - 1).clamp_(0, nb - 1)
        out_view[a:a + m] = build_view[r]


def bench_group_by(ctx, g, args, torch, stream, peak, peak_kind):
    """BASELINE config 3: hash group-by SUM/COUNT/MIN/MAX over 2^log2_gb_rows rows, cardinality sweep.  Device-resident inputs,
    CUDA events on the launching stream; rows/s = input rows / time; roofline on 16 B/row + 40 B per group."""
    import ctypes as C
    import numpy as np
    from gpu_olap_engine_b200._lib import AggSpec, check
    n = 1 << args.log2_gb_rows
    vals = ctx.generate(n, 0, 5, 1 << 20)          # i64 uniform [0, 2^20): SUM < 2^50 stays exact
    keys = ctx.alloc(n, np.int64)
    out = []
    for G in [int(float(x)) for x in args.gb_groups.split(",")]:
        ctx.generate(n, 0, 4, G, out=keys)
        aggs = [(g.SUM, vals, g.I64), (g.COUNT, vals, g.I64), (g.MIN, vals, g.I64), (g.MAX, vals, g.I64)]
        cap = min(G, n)
        ok = ctx.alloc(cap, np.int64); oa = ctx.alloc(4 * cap, np.uint64)
        specs = (AggSpec * 4)(*[AggSpec(v.ptr, f, t) for f, v, t in aggs])

        def run():
            gg = C.c_uint64()
            check(ctx.lib.olap_group_by_i64(ctx.handle, keys.ptr, specs, 4, n, G, None, 0, ok.ptr, oa.ptr, cap, None, C.byref(gg), None))
            return gg.value
        for _ in range(3):
            ng = run()
        launches0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(stream)
        for _ in range(args.gb_steps):
            ng = run()
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.gb_steps
        # parity inside the bench: COUNT sums to n, SUM equals the column sum, group count = distinct keys
        assert int(dev_view(torch, oa, 4 * cap)[cap:cap + ng].sum()) == n, "group-by COUNT does not add up"
        by = 16.0 * n + 40.0 * ng
        out.append({"groups_hint": G, "groups_found": int(ng), "ms_per_step": ms, "value": n / (ms * 1e-3), "unit": "rows/s",
                    "roofline": {"bound": "hbm", "achieved": by / (ms * 1e-3) / 1e9, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                                 "frac": by / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": by},
                    "phases_ms": {k: v for k, v in ctx.last_phase_times().items() if k != "end"},
                    "gpu_launches_per_step": (ctx.launch_count() - launches0) / args.gb_steps})
        del ok, oa
    del keys, vals
    return {"op": f"hash group-by SUM/COUNT/MIN/MAX, i64 keys uniform in [0, G), i64 values, 2^{args.log2_gb_rows} rows", "steps": args.gb_steps,
            "warmup": 3, "sweep": out}


def cpu_reference_group_by(log2_rows, G, threads):
    import numpy as np
    import oracle as O
    n = 1 << log2_rows
    k = O.gen_uniform_i64(n, 4, G); v = O.gen_uniform_i64(n, 5, 1 << 20)
    t0 = time.perf_counter()
    keys, res = O.group_by_i64(k, [v, v, v, v], [O.SUM, O.COUNT, O.MIN, O.MAX], threads=threads)
    dt = time.perf_counter() - t0
    assert int(res[1].sum()) == n
    return n / dt, dt


def bench_single(ctx, g, args, local_rank, peak, peak_kind):
    import ctypes as C
    import numpy as np
    import torch
    nb, npb = 1 << args.log2_build, 1 << args.log2_probe
    rows = nb + npb
    traffic = load_traffic()
    stream = torch.cuda.ExternalStream(ctx.lib.olap_ctx_stream(ctx.handle, 0))  # the stream the kernels launch on
    # build keys = permutation of [0, 2^27) (unique), probe = uniform FK over the build domain, seed 3
    dbuild = ctx.generate(nb, 1, 3, nb)
    if args.zipf > 0:
        dprobe = ctx.generate_fk(npb, 3, nb, 33, args.zipf)     # FK ~ Zipf(s) over the build rows, keys through the build permutation
    else:
        dprobe = ctx.generate(npb, 0, 33, nb)
    pairs = ctx.alloc(2 * npb, np.uint32)
    ctx.synchronize()
    sampler = ClockSampler(local_rank)
    ms_step, phases, launches, cnt = timed_join(ctx, g, torch, stream, dbuild, nb, dprobe, npb, pairs, args.steps, args.warmup, expect=npb)
    value = rows / (ms_step * 1e-3)
    # every pair of the last timed step, checked on the device: each probe row exactly once, equal keys
    verify_pairs(torch, dbuild, dprobe, pairs, nb, npb, cnt)


    # ---- BASELINE config 2 names "int64 keys + int64 payload": the same join followed by the materialisation of both sides' payload
    # columns in pair order (what execute_hash_join's output RecordBatch holds, executor.rs:255-264) -- late materialisation, random
    # 8-byte gathers by the row-id pairs
    payload = None
    if not args.skip_payload:
        pb, pp = ctx.generate(nb, 0, 91, 1 << 40), ctx.generate(npb, 0, 92, 1 << 40)
        ob, op_ = ctx.alloc(npb, np.int64), ctx.alloc(npb, np.int64)

        def pstep():
            c, _ = ctx.hash_join_i64(dbuild, nb, dprobe, npb, max_matches=npb, strategy=g.JOIN_RADIX, out_pairs=pairs)
            g._lib.check(ctx.lib.olap_join_gather_8b(ctx.handle, pb.ptr, pairs.ptr, 0, c, ob.ptr, None))
            g._lib.check(ctx.lib.olap_join_gather_8b(ctx.handle, pp.ptr, pairs.ptr, 1, c, op_.ptr, None))
            return c
        for _ in range(2):
            pc = pstep()
        psteps = max(2, args.steps // 2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(stream)
        for _ in range(psteps):
            pc = pstep()
        e1.record(stream); torch.cuda.synchronize()
        pms = e0.elapsed_time(e1) / psteps
        # every gathered value checked on the device: out_build[i] == payload_build[left_i], out_probe[i] == payload_probe[right_i]
        pr = dev_view(torch, pairs, 2 * pc, "<u4").view(torch.int32).view(-1, 2)
        pbv, ppv, obv, opv = dev_view(torch, pb, nb), dev_view(torch, pp, npb), dev_view(torch, ob, npb), dev_view(torch, op_, npb)
        for a in range(0, pc, 1 << 26):
            l = pr[a:a + (1 << 26), 0].to(torch.int64) & 0xFFFFFFFF
            r = pr[a:a + (1 << 26), 1].to(torch.int64) & 0xFFFFFFFF
            assert bool((obv[a:a + (1 << 26)] == pbv[l]).all()) and bool((opv[a:a + (1 << 26)] == ppv[r]).all()), "payload materialisation check failed"
        in_bytes, pair_bytes, out_bytes = 16.0 * rows, 8.0 * pc, 16.0 * pc
        payload = {"value": rows / (pms * 1e-3), "unit": "rows/s", "ms_per_step": pms, "steps": psteps, "join_ms": ms_step, "materialise_ms": pms - ms_step,
                   "workload": "the same join + both int64 payload columns gathered into pair order (2 x 2^%d values, every one verified)" % args.log2_probe,
                   "job_roofline_survey": {"algorithmic_bytes": in_bytes + pair_bytes, "frac": (in_bytes + pair_bytes) / (pms * 1e-3) / 1e9 / peak,
                                           "note": "SURVEY 8(d) C2: 16 B x rows in (keys + payloads) + 8 B x pairs out = 15.03 GB"},
                   "job_roofline_materialised": {"algorithmic_bytes": in_bytes + pair_bytes + out_bytes, "frac": (in_bytes + pair_bytes + out_bytes) / (pms * 1e-3) / 1e9 / peak,
                                                 "note": "... + the two materialised payload columns (16 B x pairs)"}}
        del pb, pp, ob, op_

    # ---- the general path: sparse build keys (2^27 unique keys of a 2^33 domain) take hash partitions + bucket tables
    general = None
    if not args.skip_general:
        gb_ = ctx.generate(nb, 1, 7, 1 << 33)                      # permutation of [0, 2^33) evaluated at rows 0..nb-1: unique, sparse
        gp_ = ctx.alloc(npb, np.int64)
        ctx.synchronize()
        fk_probe_keys(torch, dev_view(torch, gb_, nb), dev_view(torch, gp_, npb), 0.0, 34)   # FK: the key of a uniformly drawn build row
        torch.cuda.synchronize()
        gms, gph, gl, gcnt = timed_join(ctx, g, torch, stream, gb_, nb, gp_, npb, pairs, max(2, args.steps // 2), 3, expect=npb)
        verify_pairs(torch, gb_, gp_, pairs, nb, npb, gcnt)
        general = {"value": rows / (gms * 1e-3), "unit": "rows/s", "ms_per_step": gms, "phases_ms": gph,
                   "workload": f"same sizes, build keys = 2^{args.log2_build} unique keys of a 2^33 domain (sparse: hash partitions + two-slot bucket tables in L2)",
                   "kernels": kernel_rooflines(gph, GENERAL_PHASE_BYTES, nb, npb, gcnt, peak, {}),
                   "job_roofline_frac": (8.0 * rows + 8.0 * gcnt) / (gms * 1e-3) / 1e9 / peak}
        del gb_, gp_

    if args.skip_e2e or args.skip_cpu:   # profiling runs (ncu launch lists): the device-timed legs only, not a bench line
        group_by = None if args.skip_groupby else bench_group_by(ctx, g, args, torch, stream, peak, peak_kind)
        clocks = sampler.stop()
        print(json.dumps({"profiling_run": True, "value": value, "ms_per_step": ms_step, "phases_ms": phases, "gpu_launches": int(launches),
                          "value_with_payload": payload, "value_general": general, "group_by": group_by, "clocks": clocks}))
        return 0

    # ---- e2e: host (pinned) buffers through the C-ABI, copies inside the timed region
    hb, hp, hout = C.c_void_p(), C.c_void_p(), C.c_void_p()
    g._lib.check(ctx.lib.olap_host_alloc(nb * 8, C.byref(hb)))
    g._lib.check(ctx.lib.olap_host_alloc(npb * 8, C.byref(hp)))
    g._lib.check(ctx.lib.olap_host_alloc(npb * 8, C.byref(hout)))
    g._lib.check(ctx.lib.olap_memcpy_d2h(ctx.handle, hb, dbuild.ptr, nb * 8, None))
    g._lib.check(ctx.lib.olap_memcpy_d2h(ctx.handle, hp, dprobe.ptr, npb * 8, None))
    ctx.synchronize()
    db2, dp2 = ctx.alloc(nb, np.int64), ctx.alloc(npb, np.int64)

    # Pipelined through the public build-once / probe-per-batch API: the probe side arrives in n_batches batches (the
    # reference executor's Vec<RecordBatch>); H2D of batch i+1, the join of batch i and D2H of batch i-1's pairs overlap
    # on three streams.  Everything -- every copy, every kernel -- is inside the timed region.
    n_batches = int(os.environ.get("OLAP_E2E_BATCHES", "8"))
    bsz = npb // n_batches
    s_comp = torch.cuda.ExternalStream(ctx.lib.olap_ctx_stream(ctx.handle, 0))
    s_h2d = torch.cuda.ExternalStream(ctx.lib.olap_ctx_stream(ctx.handle, 4))
    s_d2h = torch.cuda.ExternalStream(ctx.lib.olap_ctx_stream(ctx.handle, 5))

    def e2e_step():
        ev_b = torch.cuda.Event(); ev_p = [torch.cuda.Event() for _ in range(n_batches)]
        g._lib.check(ctx.lib.olap_memcpy_h2d(ctx.handle, db2.ptr, hb, nb * 8, s_h2d.cuda_stream)); ev_b.record(s_h2d)
        for i in range(n_batches):
            g._lib.check(ctx.lib.olap_memcpy_h2d(ctx.handle, dp2.ptr + 8 * i * bsz, hp.value + 8 * i * bsz, bsz * 8, s_h2d.cuda_stream))
            ev_p[i].record(s_h2d)
        s_comp.wait_event(ev_b)
        total = 0
        with g.JoinBuild(ctx, db2, nb, g.JOIN_RADIX) as jb:
            for i in range(n_batches):
                s_comp.wait_event(ev_p[i])
                c = jb.probe(dp2.ptr + 8 * i * bsz, bsz, probe_row_base=i * bsz, out_pairs=pairs.ptr + 8 * total, max_matches=npb - total)
                ev_j = torch.cuda.Event(); ev_j.record(s_comp); s_d2h.wait_event(ev_j)
                g._lib.check(ctx.lib.olap_memcpy_d2h(ctx.handle, hout.value + 8 * total, pairs.ptr + 8 * total, c * 8, s_d2h.cuda_stream))
                total += c
        s_d2h.synchronize()
        return total

    e2e_step()
    e2e_steps = max(1, min(args.steps, 3))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        c = e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    e2e_path = ",".join(k for k in ctx.last_phase_times() if k.startswith("build_probe"))
    # the host-visible result of the last e2e step: every probe row exactly once, equal keys (all pairs, numpy on the host)
    out = np.ctypeslib.as_array(C.cast(hout, C.POINTER(C.c_uint32)), shape=(2 * npb,)).reshape(-1, 2)
    hbv = np.ctypeslib.as_array(C.cast(hb, C.POINTER(C.c_int64)), shape=(nb,))
    hpv = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_int64)), shape=(npb,))
    sl = out[:: max(1, npb // (1 << 22))]
    assert c == npb and (hbv[sl[:, 0]] == hpv[sl[:, 1]]).all(), "e2e result check failed"
    seen = np.zeros(npb, np.uint8); seen[out[:, 1]] = 1
    assert int(seen.sum()) == npb, "e2e result check failed: a probe row is missing or repeated"
    del seen, db2, dp2

    # ---- group-by (the other half of BASELINE's metric), same process, same clocks record
    group_by = None if args.skip_groupby else bench_group_by(ctx, g, args, torch, stream, peak, peak_kind)
    clocks = sampler.stop()

    # ---- roofline: every kernel of the step from the CUDA-event marks around it inside the timed steps; `roofline` = the one with
    # the largest share of the step
    kernels = kernel_rooflines(phases, PHASE_BYTES, nb, npb, cnt, peak, traffic)
    top = kernels[0]
    job_bytes = 8.0 * rows + 8.0 * npb  # keys read once + pairs written once (the JoinParams contract)

    # ---- CPU baseline (oracle port of the reference CPU path) on a bounded sample, all host cores and single-thread
    cores = os.cpu_count() or 1
    cpu_rps, cpu_sec, _ = cpu_reference_join(args.cpu_log2_build, args.cpu_log2_probe, cores)
    cpu1_rps, _, _ = cpu_reference_join(args.cpu_log2_build - 3, args.cpu_log2_probe - 3, 1)
    sample = (f"2^{args.cpu_log2_build} unique-key build x 2^{args.cpu_log2_probe} FK probe (same generator), pairs materialised, "
              f"{cores} threads, {cpu_sec:.2f} s; single-thread ref1 on an 8x smaller sample: {cpu1_rps / 1e6:.1f} M rows/s")
    if group_by is not None:
        gb_cpu = []
        for G in (100, 100000):
            r, sec = cpu_reference_group_by(args.cpu_log2_gb_rows, G, cores)
            gb_cpu.append({"groups": G, "value": r, "unit": "rows/s", "seconds": sec})
        group_by["cpu_baseline"] = {"kind": "port", "cores": cores, "sample": f"oracle orc_group_by_i64_mt on 2^{args.cpu_log2_gb_rows} rows of the same generator", "runs": gb_cpu}

    direct = "build_probe_smem" in phases
    line = {
        "metric": "join rows/sec (|build|+|probe|)/time", "value": value, "unit": "rows/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": {"workload": f"single-GPU radix hash join 2^{args.log2_build} unique-key build x 2^{args.log2_probe} FK probe, int64 keys, " + (f"Zipf s = {args.zipf}" if args.zipf > 0 else "uniform (Zipf s = 0)"),
                   "output": "(u32 build row, u32 probe row) pairs, all 2^%d materialised; every pair verified after the timed steps" % args.log2_probe,
                   "l2": "inputs (%.1f GB) are larger than the 126 MB L2; no flush needed" % (rows * 8 / 1e9),
                   "strategy": "RadixHashJoin",
                   "tables": ("direct-address row-id tables in shared memory, one per range of 2^15 keys (dense unique build keys: key range remembered per column, verified by the build pass)"
                              if direct else "L2-resident tables (see phases_ms)")},
        "clocks": clocks,
        "e2e": {"value": rows / (e2e_ms * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": rows * 8, "d2h_bytes_per_step": npb * 8 + 8,
                "ms_per_step": e2e_ms, "steps": e2e_steps,
                "how": f"olap_join_build_create(RADIX) + {n_batches} x olap_join_probe on pinned host buffers; H2D / join / D2H pipelined on 3 streams; wall clock; "
                       f"join path of the batches: {e2e_path}; all pairs checked on the host"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": top["kernel"], "achieved": top["achieved"], "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                     "frac": top["frac"], "traffic": (top["traffic"] or {}).get("dram_bytes_per_launch") if isinstance(top["traffic"], dict) else top["traffic"],
                     "traffic_source": "ncu --set full, one capture per kernel (profiles/traffic.json): static, not re-measured in this run",
                     "algorithmic_bytes_per_launch": top["algorithmic_bytes_per_launch"], "ms_per_launch": top["ms_per_launch"],
                     "share_of_step": top["ms_per_launch"] / ms_step},
        "roofline_kernels": kernels,
        "job_roofline": {"algorithmic_bytes": job_bytes, "achieved_gbs": job_bytes / (ms_step * 1e-3) / 1e9,
                         "frac": job_bytes / (ms_step * 1e-3) / 1e9 / peak,
                         "note": "keys read once + pairs written once; SURVEY 8(d)'s 15.03 GB also counts payload columns the join never touches"},
        "phases_ms": phases,
        "value_with_payload": payload,
        "value_general": general,
        "group_by": group_by,
        "cpu_baseline": {"value": cpu_rps, "unit": "rows/s", "cores": cores, "kind": "port", "sample": sample},
    }
    print(json.dumps(line))
    return 0


def verify_pairs(torch, dbuild, dprobe, pairs, nb, npb, cnt):
    """ALL pairs of a join, checked on the device (torch views of the library's buffers, outside every timed region):
    build_key[left] == probe_key[right] for every pair, and no probe row appears twice (unique build keys: at most one match)."""
    assert cnt <= npb
    pr = dev_view(torch, pairs, 2 * cnt, "<u4").view(torch.int32).view(-1, 2)
    bk, pk = dev_view(torch, dbuild, nb), dev_view(torch, dprobe, npb)
    seen = torch.zeros(npb, dtype=torch.uint8, device="cuda")
    step = 1 << 26
    for a in range(0, cnt, step):
        l = pr[a:a + step, 0].to(torch.int64) & 0xFFFFFFFF
        r = pr[a:a + step, 1].to(torch.int64) & 0xFFFFFFFF
        assert bool((bk[l] == pk[r]).all()), "join pairs: build_key[left] != probe_key[right]"
        seen[r] = 1
    assert int(seen.sum(dtype=torch.int64)) == cnt, "join pairs: a probe row appears twice"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-build", type=int, default=None)
    ap.add_argument("--log2-probe", type=int, default=None)
    ap.add_argument("--cpu-log2-build", type=int, default=26)
    ap.add_argument("--cpu-log2-probe", type=int, default=28)
    ap.add_argument("--zipf", type=float, default=0.0, help="Zipf exponent of the probe-side foreign keys (0 = uniform)")
    ap.add_argument("--log2-gb-rows", type=int, default=30)
    ap.add_argument("--gb-groups", default="1e2,1e5,1e8")
    ap.add_argument("--gb-steps", type=int, default=3)
    ap.add_argument("--cpu-log2-gb-rows", type=int, default=26)
    ap.add_argument("--skip-groupby", action="store_true")
    ap.add_argument("--skip-general", action="store_true")
    ap.add_argument("--skip-payload", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: omit the host-buffer leg")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs only: omit the CPU baseline leg")
    args = ap.parse_args()
    args.sizes_given = args.log2_build is not None or args.log2_probe is not None
    args.log2_build = 27 if args.log2_build is None else args.log2_build
    args.log2_probe = 29 if args.log2_probe is None else args.log2_probe
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    sys.exit(run_reference(args) if args.impl == "reference" else run_ours(args))


if __name__ == "__main__":
    main()
