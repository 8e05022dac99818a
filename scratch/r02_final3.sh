mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29511 tests/run_sharded_nccl.py > gpurun_out/sharded_nccl_w2.log 2>&1; echo "nccl test rc=$?"; grep "ok:\|Error\|error\|assert" gpurun_out/sharded_nccl_w2.log | tail -12
for z in 0.5 1.0; do
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --skip-e2e --skip-cpu --zipf $z > gpurun_out/r02_bench_n2_zipf${z}_final.json 2> gpurun_out/bench_n2_zipf$z.err; echo "zipf $z rc=$?"
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n2_zipf$z.err | tail -3
python - gpurun_out/r02_bench_n2_zipf${z}_final.json <<'PY'
import sys, json
l = json.loads([x for x in open(sys.argv[1]).read().splitlines() if x.startswith('{"metric')][-1]); print('step %.2f ms %.1f Grows/s' % (l['ms_per_step'], l['value'] / 1e9))
PY
done
