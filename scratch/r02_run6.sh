mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_operators.py tests/test_gpu_kernel_seam.py tests/test_gpu_shard.py tests/test_gpu_djoin.py -q -x --timeout 600 > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_new.log
timeout 120 python scratch/shard_zipf_prof.py 26 28 2 2 1.0 > gpurun_out/zipf_prof_after.log 2>&1; echo "zipf rc=$?"; tail -3 gpurun_out/zipf_prof_after.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/zipf_launches_after.csv python scratch/shard_zipf_prof.py 26 28 2 1 1.0 > /dev/null 2>&1; grep "rpart_kernel<0, 1" gpurun_out/zipf_launches_after.csv | cut -d, -f5,13- | head -4
