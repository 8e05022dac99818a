"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/olap_b200.h declares (no
compute calls without a GPU), host-only entry points answer, and the reference's boundary constants hold."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "olap_b200.h")).read()
    return sorted(set(re.findall(r"OLAP_API[^;(]*?\b(olap_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import ctypes
    import gpu_olap_engine_b200 as g
    lib = ctypes.CDLL(g.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 45
    for s in syms:
        getattr(lib, s)                       # AttributeError if the header declares something the library lacks
    assert set(syms) == set(g._lib.SIGNATURES), set(syms) ^ set(g._lib.SIGNATURES)


def test_reference_boundary_constants():
    from gpu_olap_engine_b200.kernels import AggFunc, CmpOp, FilterKernel, JoinKernel, JoinStrategy, compute_launch_params
    # join_kernel.rs:85-87
    assert JoinKernel.choose_strategy(500_000) == JoinStrategy.BroadcastHashJoin
    assert JoinKernel.choose_strategy(5_000_000) == JoinStrategy.RadixHashJoin
    assert JoinKernel.choose_strategy(1_000_000) == JoinStrategy.BroadcastHashJoin and JoinKernel.choose_strategy(1_000_001) == JoinStrategy.RadixHashJoin
    # aggregate_kernel.rs:139-143
    assert (AggFunc.Sum.identity_value(), AggFunc.Min.identity_value(), AggFunc.Max.identity_value()) == (0, 2**63 - 1, -2**63)
    assert AggFunc.Count.identity_value() == 0 and AggFunc.Avg.identity_value() == 0
    # filter_kernel.rs:87-95
    assert compute_launch_params(1024) == (4, 256) and FilterKernel.compute_launch_params(1) == (1, 256)
    assert compute_launch_params(2**33)[0] == 2**33 // 256      # no `n_rows as u32` truncation (Appendix B #8)
    assert [int(c) for c in CmpOp] == [0, 1, 2, 3, 4, 5]        # codegen.rs:10


def test_engine_config_defaults():
    import gpu_olap_engine_b200 as g
    c = g.EngineConfig()                                         # lib.rs:33-43
    assert (c.max_gpu_memory, c.num_streams, c.use_unified_memory, c.batch_size, c.enable_cache) == (8 << 30, 8, True, 1 << 20, True)


def test_host_side_hash_helpers_and_plan_builders():
    import gpu_olap_engine_b200 as g
    L = g.load()
    parts = [L.olap_partition_of(k, 8) for k in range(10000)]
    assert min(parts) >= 0 and max(parts) <= 255 and len(set(parts)) == 256
    dest = [L.olap_shuffle_dest_of(k, 3) for k in range(10000)]
    assert set(dest) == set(range(8))
    agree = sum(1 for k in range(10000) if L.olap_shuffle_dest_of(k, 3) == L.olap_partition_of(k, 3))
    assert 800 < agree < 1700                                    # destination bits are independent of the local radix bits
    e = L.olap_expr_binary(L.olap_expr_column(0, b"a"), b">", L.olap_expr_literal(b"10"))
    p = L.olap_plan_filter(L.olap_plan_table_scan(b"t", None, None, -1), e)
    assert p
    L.olap_plan_free(p)
    assert not L.olap_plan_filter(None, None)


def test_context_creation_fails_loudly_without_gpu():
    import torch
    import gpu_olap_engine_b200 as g
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(g.OlapError):
        g.Context(0)


def test_oracle_is_not_imported_by_the_product():
    pkg = os.path.join(ROOT, "gpu-olap-engine_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_ctypes_structures_match_the_header_layout(tmp_path):
    # the POD structs that cross the C ABI: sizes and field offsets as gcc sees them in include/olap_b200.h must equal the
    # ctypes mirrors the tests and the bench use
    import ctypes as C
    import subprocess
    from gpu_olap_engine_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "layout.c"
    src.write_text(r"""
#include <stddef.h>
#include <stdio.h>
#include "olap_b200.h"
int main(void) {
  printf("olap_filter_term %zu %zu %zu %zu %zu %zu\n", sizeof(olap_filter_term), offsetof(olap_filter_term, lhs), offsetof(olap_filter_term, rhs),
         offsetof(olap_filter_term, literal_bits), offsetof(olap_filter_term, op), offsetof(olap_filter_term, type));
  printf("olap_agg_spec %zu %zu %zu %zu\n", sizeof(olap_agg_spec), offsetof(olap_agg_spec, values), offsetof(olap_agg_spec, func),
         offsetof(olap_agg_spec, value_type));
  printf("olap_expr_prog %zu %zu %zu %zu %zu\n", sizeof(olap_expr_prog), offsetof(olap_expr_prog, n_ops), offsetof(olap_expr_prog, op),
         offsetof(olap_expr_prog, column), offsetof(olap_expr_prog, literal_bits));
  printf("olap_pair %zu %d\n", sizeof(olap_pair), OLAP_PROG_MAX_OPS);
  return 0;
}
""")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    got = {ln.split()[0]: [int(x) for x in ln.split()[1:]] for ln in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines()}
    ft, ag, pr = _lib.FilterTerm, _lib.AggSpec, _lib.ExprProg
    assert got["olap_filter_term"] == [C.sizeof(ft), ft.lhs.offset, ft.rhs.offset, ft.literal_bits.offset, ft.op.offset, ft.type.offset]
    assert got["olap_agg_spec"] == [C.sizeof(ag), ag.values.offset, ag.func.offset, ag.value_type.offset]
    assert got["olap_expr_prog"] == [C.sizeof(pr), pr.n_ops.offset, pr.op.offset, pr.column.offset, pr.literal_bits.offset]
    assert got["olap_pair"] == [8, _lib.PROG_MAX_OPS]



def test_header_is_plain_c_and_declares_the_multi_gpu_group(tmp_path):
    # the drop-in boundary is a C ABI: the header must compile as C99 (and as C++) on its own, and carry SURVEY 8(b)'s multi-GPU group
    import shutil
    import subprocess
    src = tmp_path / "h.c"
    src.write_text('#include "olap_b200.h"\nint main(void) { return 0; }\n')
    inc = os.path.join(ROOT, "include")
    if shutil.which("gcc"):
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)])
    if shutil.which("g++"):
        subprocess.check_call(["g++", "-std=c++11", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)])
    syms = set(_declared_symbols())
    assert {"olap_shard_ctx_create_all", "olap_sharded_hash_join_i64", "olap_sharded_group_by_i64", "olap_shard_route_i64",
            "olap_join_unmatched_rows", "olap_join_gather_8b", "olap_group_by_expr_i64", "olap_register_device_batch"} <= syms
