# where does a Zipf-skewed sharded join spend its step?  stream traces of rank 0 at N=2, uniform vs Zipf 1.0
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for z in 0 1.0; do
OLAP_SHARD_TRACE=1 timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 3 --skip-e2e --skip-cpu --zipf $z > gpurun_out/trace_n2_zipf$z.log 2>&1; echo "zipf $z rc=$?"
grep "shard trace" gpurun_out/trace_n2_zipf$z.log | tail -1 | cut -c1-2500
done
