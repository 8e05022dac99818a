# round-2 1-GPU measurement pass: parity suite, the bench line, the ncu launch list of the same command, one ncu --set full capture
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log ) 
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_n1.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu --gb-steps 1 > gpurun_out/bench_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"rpart_kernel|djoin_kernel" -s 15 -c 5 -o gpurun_out/r02_join python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu --skip-general --skip-groupby > gpurun_out/bench_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gr_part_kernel|gr_agg_kernel" -c 2 -o gpurun_out/r02_groupby python bench_ops.py --only groupby --groups 100000 --steps 1 --warmup 0 > gpurun_out/bench_ncu3.log 2>&1; echo "ncu3 rc=$?"
timeout 200 python bench_ops.py --only groupby,filter > gpurun_out/r02_bench_ops.jsonl 2> gpurun_out/bench_ops.err; echo "ops rc=$?"
ls -la gpurun_out
