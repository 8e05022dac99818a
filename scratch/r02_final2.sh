# final 2-GPU pass of round 2: the bench line at N=2 (per GPU 2^28 x 2^30), then Zipf-skewed probe keys
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n2.err | tail -5
for z in 0.5 1.0 1.5; do
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --skip-e2e --skip-cpu --zipf $z > gpurun_out/r02_bench_n2_zipf${z}_final.json 2> gpurun_out/bench_n2_zipf$z.err; echo "zipf $z rc=$?"
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n2_zipf$z.err | tail -3
done
