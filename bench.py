#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 execution core (contract: see the task prompt / DESIGN.md "Measurement").

Metric (BASELINE.json): join rows/sec, rows = |build| + |probe|, on the single-GPU radix hash join
2^27 unique-key build x 2^29 FK probe, int64 keys (configs[1]); % of HBM roofline for the dominant kernel.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2-build B --log2-probe P]

* value     device-resident inputs, CUDA events around K joins on the launching stream, max over ranks
* e2e       the same join through the C-ABI with HOST (pinned) buffers: H2D of both key columns, join, D2H of the
            count and of the row-id pairs, all inside the timed region
* roofline  dominant kernel (join_pipeline_kernel: one persistent launch per step): algorithmic bytes / its CUDA-event
            duration in the timed region, against MEASURED_PEAKS.json's hbm_gbs; the partition kernel next to it
* cpu_baseline / --impl reference   the oracle's multi-threaded port of the reference CPU path on this box's cores
N > 1: weak scaling, one process per GPU (torchrun); the join is key-sharded with an NCCL all-to-all shuffle
(gpu_olap_engine_b200/sharded.py); value = rows of all ranks / max-over-ranks time.
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

    ctx = g.Context(local_rank, 0, 8)
    ctx.set_profiling(True)
    peak, peak_kind = peak_hbm()

    if world > 1:
        from gpu_olap_engine_b200 import sharded
        return sharded.bench_sharded_join(ctx, args, rank, world, local_rank, peak, peak_kind, ClockSampler)

    nb, npb = 1 << args.log2_build, 1 << args.log2_probe
    rows = nb + npb
    # build keys = permutation of [0, 2^27) (unique), probe = uniform FK over the build domain, seed 3
    dbuild = ctx.generate(nb, 1, 3, nb)
    dprobe = ctx.generate(npb, 0, 33, nb)
    pairs = ctx.alloc(2 * npb, np.uint32)
    ctx.synchronize()

    def step():
        c, _ = ctx.hash_join_i64(dbuild, nb, dprobe, npb, max_matches=npb, strategy=g.JOIN_RADIX, out_pairs=pairs)
        return c

    for _ in range(args.warmup):
        cnt = step()
    assert cnt == npb or os.environ.get("OLAP_JOIN_DEBUG"), f"join count {cnt} != {npb}"
    sampler = ClockSampler(local_rank)
    stream = torch.cuda.ExternalStream(ctx.lib.olap_ctx_stream(ctx.handle, 0))  # the stream the kernels launch on
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase_acc = {}
    launches0 = ctx.launch_count()
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
        for k, v in ctx.last_phase_times().items():
            phase_acc[k] = phase_acc.get(k, 0.0) + v
    ev1.record(stream)
    torch.cuda.synchronize()
    launches = ctx.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    ms_step = ms_total / args.steps
    value = rows / (ms_step * 1e-3)
    phases = {k: v / args.steps for k, v in phase_acc.items()}

    if args.skip_e2e or args.skip_cpu:
        clocks = sampler.stop()
        print(json.dumps({"profiling_run": True, "value": value, "ms_per_step": ms_step, "phases_ms": phases, "gpu_launches": int(launches),
                          "clocks": clocks}))
        return 0
    # ---- e2e: host (pinned) buffers through the C-ABI, copies inside the timed region
    import ctypes as C
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
    e2e_strategy = g.JOIN_RADIX if os.environ.get("OLAP_E2E_STRATEGY") == "radix" else g.JOIN_BROADCAST
    # batch streaming uses the single-table strategy: built once, nothing is rebuilt per batch (measured device-resident:
    # 34 ms for 8 batches vs 59 ms when every batch re-builds the 256 L2-sized partition tables)
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
        with g.JoinBuild(ctx, db2, nb, e2e_strategy) as jb:
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
    clocks = sampler.stop()
    # spot-check the host-visible result of the last e2e step
    out = np.ctypeslib.as_array(C.cast(hout, C.POINTER(C.c_uint32)), shape=(2 * npb,)).reshape(-1, 2)
    hbv = np.ctypeslib.as_array(C.cast(hb, C.POINTER(C.c_int64)), shape=(nb,))
    hpv = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_int64)), shape=(npb,))
    sl = out[:: max(1, npb // 100000)]
    assert c == npb and (hbv[sl[:, 0]] == hpv[sl[:, 1]]).all(), "e2e result check failed"

    # ---- roofline.  Dominant kernel by share of the step (profiles/r01_ncu_launches_*_summary.csv): join_pipeline_kernel,
    # ONE persistent launch per step that clears, builds and probes the L2-resident tables of all partitions.  Algorithmic
    # bytes per launch = partitioned (key 8 B + row id 4 B) of both sides read once + 8 B per pair written.  Its duration is
    # the CUDA-event time between the marks around that launch inside the timed steps.  The second kernel (probe-side
    # partition_kernel: reads 8 B key, writes 8 B key + 4 B row id per row) is reported next to it.
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass
    direct = "build_probe_direct" in phases   # dense unique build keys detected: direct-address tables instead of bucket tables
    bp_ms = phases.get("build_probe_direct") or phases.get("build_probe")
    bp_bytes = 12.0 * rows + 8.0 * npb
    bp_gbs = bp_bytes / (bp_ms * 1e-3) / 1e9 if bp_ms else None
    part_ms = phases.get("partition_probe")
    part_bytes = 20.0 * npb
    achieved = part_bytes / (part_ms * 1e-3) / 1e9 if part_ms else None
    job_bytes = 8.0 * rows + 8.0 * npb  # keys read once + pairs written once (the JoinParams contract)

    # ---- CPU baseline (oracle port of the reference CPU path) on a bounded sample, all host cores and single-thread
    cores = os.cpu_count() or 1
    cpu_rps, cpu_sec, _ = cpu_reference_join(args.cpu_log2_build, args.cpu_log2_probe, cores)
    cpu1_rps, _, _ = cpu_reference_join(args.cpu_log2_build - 3, args.cpu_log2_probe - 3, 1)
    sample = (f"2^{args.cpu_log2_build} unique-key build x 2^{args.cpu_log2_probe} FK probe (same generator), pairs materialised, "
              f"{cores} threads, {cpu_sec:.2f} s; single-thread ref1 on an 8x smaller sample: {cpu1_rps / 1e6:.1f} M rows/s")

    line = {
        "metric": "join rows/sec (|build|+|probe|)/time", "value": value, "unit": "rows/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": {"workload": f"single-GPU radix hash join 2^{args.log2_build} unique-key build x 2^{args.log2_probe} FK probe, int64 keys, uniform",
                   "output": "(u32 build row, u32 probe row) pairs, all 2^%d materialised" % args.log2_probe,
                   "l2": "inputs (%.1f GB) are larger than the 126 MB L2; no flush needed" % (rows * 8 / 1e9),
                   "strategy": "RadixHashJoin",
                   "tables": ("direct-address row-id tables per key range (dense, unique build keys detected at run time)" if direct
                              else "two-slot bucket tables per hash partition")},
        "clocks": clocks,
        "e2e": {"value": rows / (e2e_ms * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": rows * 8, "d2h_bytes_per_step": npb * 8 + 8,
                "ms_per_step": e2e_ms, "steps": e2e_steps,
                "how": "olap_join_build_create + 8 x olap_join_probe on pinned host buffers; H2D / join / D2H pipelined on 3 streams; wall clock"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "join_pipeline_kernel<%s> (one persistent launch per step: clear + build + probe of all partitions' L2-resident tables)" % ("direct" if direct else "buckets"),
                     "achieved": bp_gbs, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": (bp_gbs / peak) if bp_gbs else None,
                     "traffic": traffic.get("join_pipeline_dram_bytes_per_launch" if direct else "join_pipeline_buckets_dram_bytes_per_launch"), "algorithmic_bytes_per_launch": bp_bytes, "ms_per_launch": bp_ms,
                     "note": "bound by L2 sector requests (one table gather per probe row, one atomic per build row), not by HBM: DESIGN.md section 4"},
        "roofline_partition": {"bound": "hbm", "kernel": "partition_kernel<0,true,true> (probe side; second largest launch)",
                               "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                               "traffic": traffic.get("partition_probe_dram_bytes_per_launch"), "algorithmic_bytes_per_launch": part_bytes,
                               "ms_per_launch": part_ms},
        "job_roofline": {"algorithmic_bytes": job_bytes, "achieved_gbs": job_bytes / (ms_step * 1e-3) / 1e9,
                         "frac": job_bytes / (ms_step * 1e-3) / 1e9 / peak,
                         "note": "keys read once + pairs written once; SURVEY 8(d)'s 15.03 GB also counts payload columns the join never touches"},
        "phases_ms": phases,
        "cpu_baseline": {"value": cpu_rps, "unit": "rows/s", "cores": cores, "kind": "port", "sample": sample},
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-build", type=int, default=27)
    ap.add_argument("--log2-probe", type=int, default=29)
    ap.add_argument("--cpu-log2-build", type=int, default=26)
    ap.add_argument("--cpu-log2-probe", type=int, default=28)
    ap.add_argument("--zipf", type=float, default=0.0)
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: omit the host-buffer leg")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs only: omit the CPU baseline leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    sys.exit(run_reference(args) if args.impl == "reference" else run_ours(args))


if __name__ == "__main__":
    main()
