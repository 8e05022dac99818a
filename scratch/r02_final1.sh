# final 1-GPU pass of round 2: the whole GPU suite, the bench line, the ncu launch list of the bench's profiling mode
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log )
tail -6 gpurun_out/pytest_gpu_final.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/bench_n1_final.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench_n1_final.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_ncu_launches_final.csv python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu --gb-steps 1 > gpurun_out/bench_ncu_final.log 2>&1; echo "ncu rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
