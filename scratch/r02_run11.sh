mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gb_shared_kernel -s 1 -c 1 -o gpurun_out/r02_q1 python bench_ops.py --only tpch --sf 10 --steps 1 --warmup 1 > gpurun_out/tpch_ncu2.log 2>&1; echo "ncu2 rc=$?"
