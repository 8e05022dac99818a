N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 400 gpurun_out/bench_n$N.err | grep -v "^W"
summ() { python - "$1" "$2" <<'PY'
import sys, json
try:
    l = json.loads([x for x in open(sys.argv[1]).read().strip().splitlines() if x.startswith('{"metric')][-1])
    o = l['overlap']
    print('RESULT %s | step %.2f ms %.1f Grows/s | shuffle %.2f ms %.1f GB/s | local %.2f | hidden %.2f | e2e %.1f Grows/s' % (sys.argv[2], l['ms_per_step'], l['value'] / 1e9, o['shuffle_alone_ms'], l['nvlink']['achieved_gbs_per_direction'], o['local_alone_ms'], o['hidden_ms'], l['e2e']['value'] / 1e9))
except Exception as e:
    print('FAILED', sys.argv[2], e); print(open(sys.argv[1]).read()[-800:])
PY
}
summ gpurun_out/r02_bench_n$N.json "N=$N uniform ce"
for z in 1.0 1.5; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --skip-e2e --skip-cpu --zipf $z > gpurun_out/r02_bench_n${N}_zipf$z.json 2> gpurun_out/bench_n${N}_zipf.err
summ gpurun_out/r02_bench_n${N}_zipf$z.json "N=$N zipf $z ce"
tail -c 1500 gpurun_out/bench_n${N}_zipf.err | grep -E "Error|flagged" | tail -2
done
OLAP_SHARD_MODE=tma OLAP_SHARD_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/r02_bench_n${N}_tma.json 2> gpurun_out/bench_n${N}_tma.err
summ gpurun_out/r02_bench_n${N}_tma.json "N=$N uniform tma"
