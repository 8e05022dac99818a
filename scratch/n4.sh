N=$1
python -m pytest tests/test_gpu_shard.py tests/test_gpu_djoin.py -x -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/run_sharded_nccl.py 2>&1 | grep -v "^W\|^\*\*\*" | grep "ok:\|Error\|error" | tail -9
run() { env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 4 --warmup 3 --skip-e2e --skip-cpu $EXTRA > /tmp/out.log 2>&1; grep '^{"metric' /tmp/out.log | tail -1 > /tmp/line.json; python - "$* $EXTRA" <<'PY'
import sys, json
try:
    l = json.loads(open('/tmp/line.json').readline())
    o = l['overlap']
    print('RESULT', sys.argv[1], '| step', round(l['ms_per_step'], 2), 'ms', round(l['value'] / 1e9, 1), 'Grows/s | shuffle', round(o['shuffle_alone_ms'], 2), 'ms', round(l['nvlink']['achieved_gbs_per_direction'], 1), 'GB/s | local', round(o['local_alone_ms'], 2), '| hidden', round(o['hidden_ms'], 2))
except Exception as e:
    print('FAILED', sys.argv[1], e); print(open('/tmp/out.log').read()[-1500:])
PY
}
EXTRA=""
run OLAP_SHARD_CHUNKS=8
EXTRA="--zipf 1.0"
run OLAP_SHARD_CHUNKS=8
