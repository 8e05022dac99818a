"""GPU tests of the sharded operators: world 1 in-process; world 2 over NCCL when two GPUs are visible."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


def test_sharded_ops_world1(ctx):
    import torch
    from gpu_olap_engine_b200 import sharded
    ops = sharded.DeviceOps(ctx)
    dev = ops.device
    b = O.gen_uniform_i64(50000, 11, 40000); p = O.gen_uniform_i64(200000, 12, 40000)
    res = sharded.ShardedJoin(ops).join(torch.from_numpy(b).to(dev), len(b), torch.from_numpy(p).to(dev), len(p), max_matches=600000)
    pairs = sharded.ShardedJoin.global_pairs(res, [0], [0]).cpu().numpy()
    ec, ep = O.hash_join_i64(b, p)
    assert res["count"] == ec and (O.sort_pairs(pairs.astype(np.uint32)) == ep).all()
    k = O.gen_uniform_i64(100000, 4, 700); v = O.gen_uniform_i64(100000, 5, 1 << 20)
    funcs = [O.SUM, O.COUNT, O.MIN, O.MAX, O.AVG]
    m, mk, out = sharded.ShardedGroupBy(ops).group_by(torch.from_numpy(k).to(dev), len(k), [torch.from_numpy(v).to(dev)] * 5, funcs, [0] * 5, 700)
    ek, eres = O.group_by_i64(k, [v] * 5, funcs)
    order = np.argsort(mk.cpu().numpy())
    assert m == len(ek) and (mk.cpu().numpy()[order] == ek).all()
    for got, exp in zip(out, eres):
        assert (got.cpu().numpy()[order] == exp).all()


def test_sharded_world2_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "run_sharded_nccl.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", script], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "sharded join ok" in r.stdout and "sharded group-by ok" in r.stdout and "range-sharded join ok" in r.stdout and "range-sharded zipf join ok" in r.stdout
