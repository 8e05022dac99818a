set -x
python scratch/shard_prof.py 25 27 4 3 > gpurun_out/shard_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"shuffle_kernel|rpart_kernel|djoin" -s 14 -c 10 -o gpurun_out/r02_shard python scratch/shard_prof.py 25 27 4 2 > gpurun_out/shard_prof_ncu.log 2>&1
tail -3 gpurun_out/shard_prof_plain.log
python bench.py --steps 2 --warmup 3 --skip-e2e > gpurun_out/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py --steps 2 --warmup 3 --skip-e2e > gpurun_out/bench_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rpart_kernel|djoin" -s 15 -c 5 -o gpurun_out/r02_join python bench.py --steps 2 --warmup 3 --skip-e2e > gpurun_out/bench_ncu2.log 2>&1
tail -2 gpurun_out/bench_plain.log
