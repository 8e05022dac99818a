python -m pytest tests/test_gpu_shard.py -x -q 2>&1 | tail -4
OLAP_SHARD_MODE=tma python -m pytest tests/test_gpu_shard.py -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_sharded_nccl.py 2>&1 | grep -v "^W\|^\*\*\*" | grep "ok:\|Error\|error" | tail -12
run() { env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 --skip-e2e --skip-cpu $EXTRA > /tmp/out.log 2>&1; grep -v "^W\|^\*\*\*" /tmp/out.log | grep "shard trace" | tail -1 | cut -c1-1800; grep '^{"metric' /tmp/out.log | tail -1 > /tmp/line.json; python - "$* $EXTRA" <<'PY'
import sys, json
try:
    l = json.loads(open('/tmp/line.json').readline())
    o = l['overlap']
    print('RESULT', sys.argv[1], '| step', round(l['ms_per_step'], 2), 'ms', round(l['value'] / 1e9, 1), 'Grows/s | shuffle', round(o['shuffle_alone_ms'], 2), 'ms', round(l['nvlink']['achieved_gbs_per_direction'], 1), 'GB/s | local', round(o['local_alone_ms'], 2), '| hidden', round(o['hidden_ms'], 2))
except Exception as e:
    print('FAILED', sys.argv[1], e); print(open('/tmp/out.log').read()[-2500:])
PY
}
EXTRA=""
run OLAP_SHARD_CHUNKS=8 OLAP_SHARD_TRACE=1
run OLAP_SHARD_CHUNKS=4
run OLAP_SHARD_CHUNKS=8 OLAP_SHARD_JOIN_GROUPS=3
EXTRA="--zipf 1.0"
run OLAP_SHARD_CHUNKS=8
