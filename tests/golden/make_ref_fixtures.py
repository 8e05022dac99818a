"""Generates tests/golden/ref_kernels.json: outputs of the REFERENCE's own CUDA kernels
(gpu-kernel-compiler/kernels/join_kernels.cuh, compiled from where it lies into oracle/_ref/libref_kernels.so by
oracle/Makefile `ref`) on small seeded inputs.  Needs a GPU; run once on a GPU box:

    gpurun -- 'python tests/golden/make_ref_fixtures.py > gpurun_out/ref_kernels.json'

and commit the file.  The CPU suite (tests/test_oracle_golden.py) then pins the oracle to what the reference actually
computed, without a GPU and without /root/reference.  Inputs are regenerated from the seeds by the oracle's generators.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from oracle import refgpu  # noqa: E402


def main():
    out = {"source": "reference kernels join_kernels.cuh:26-41 (hashes), :45-76 (radix_partition_kernel), :80-166 (build/probe), "
                     ":216-240 (hash_aggregate_kernel) executed on a B200 through oracle/ref_driver.cu"}
    k64 = [0, 1, 2, 42, -1, 2**32, 2**32 + 1, 0x0123456789abcdef, -2**63, 2**63 - 1, 1000000] + \
          [int(x) for x in O.gen_uniform_i64(64, 101, 2**63) - 2**62]
    k32 = [0, 1, 2, 42, -1, 0x7fffffff, -2**31, 123456789] + [int(x) for x in (O.gen_uniform_i64(32, 102, 2**32) - 2**31)]
    o64, o32 = refgpu.hash_vectors(k64, k32)
    out["hash_int64"] = {"keys": k64, "hashes": [int(h) for h in o64]}
    out["hash_int32"] = {"keys": k32, "hashes": [int(h) for h in o32]}
    # partition sizes of dense keys 0 .. 2^20 - 1 and of a seeded random column (radix_shift 0, 256 partitions)
    dense = np.arange(1 << 20, dtype=np.int64)
    rb, _, _ = refgpu.radix_partition(dense)
    out["partition_dense_2p20"] = {"begin": [int(x) for x in rb]}
    rnd = O.gen_uniform_i64(50000, 103, 2**40)
    rb, rk, ri = refgpu.radix_partition(rnd)
    part_of_row = np.empty(len(rnd), np.int32)
    for p in range(256):
        part_of_row[ri[rb[p]:rb[p + 1]]] = p
    out["partition_random"] = {"seed": 103, "n": 50000, "domain": 2**40, "begin": [int(x) for x in rb],
                               "partition_of_first_200_rows": [int(x) for x in part_of_row[:200]]}
    # joins: exact count and the sorted pair set, whole-table and radix flow; silent truncation
    joins = []
    for nb, npb, dom, radix in [(300, 900, 100, False), (300, 900, 100, True), (2000, 1500, 2**62, True), (64, 64, 1, False)]:
        b = O.gen_uniform_i64(nb, 111, dom); p = O.gen_uniform_i64(npb, 112, dom)
        cap = nb * npb if dom <= 100 else 10000
        c, pairs = refgpu.hash_join(b, p, cap, radix=radix)
        sp = O.sort_pairs(pairs)
        joins.append({"nb": nb, "np": npb, "domain": dom, "radix": radix, "seed_build": 111, "seed_probe": 112, "count": int(c),
                      "pairs_checksum": [int(sp[:, 0].astype(np.uint64).sum()), int(sp[:, 1].astype(np.uint64).sum()),
                                         int((sp[:, 0].astype(np.uint64) * 1000003 + sp[:, 1].astype(np.uint64)).sum() % (1 << 61))],
                      "first_pairs": [[int(a), int(bb)] for a, bb in sp[:20]]})
    out["joins"] = joins
    c, pairs = refgpu.hash_join(np.arange(1000, dtype=np.int64), np.arange(1000, dtype=np.int64), 10)
    out["join_truncation"] = {"count": int(c), "pairs_written": int(len(pairs))}
    gk = O.gen_uniform_i64(20000, 121, 50).astype(np.uint32); v = O.gen_uniform_i64(20000, 122, 1 << 20).astype(np.uint64)
    out["hash_aggregate_sum"] = {"seed_keys": 121, "seed_values": 122, "n": 20000, "groups": 50,
                                 "sums": [int(x) for x in refgpu.hash_aggregate_u64(gk, v, 50)]}
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
