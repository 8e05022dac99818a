mkdir -p gpurun_out
timeout 120 python scratch/shard_zipf_prof.py 26 28 2 2 1.0 > gpurun_out/zipf_prof_plain.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/zipf_prof_plain.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/zipf_launches.csv python scratch/shard_zipf_prof.py 26 28 2 1 1.0 > gpurun_out/zipf_prof_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"rpart_kernel<0, 1" -c 1 -o gpurun_out/r02_zipf_rpart python scratch/shard_zipf_prof.py 26 28 2 1 1.0 > gpurun_out/zipf_prof_ncu2.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out | tail -5
