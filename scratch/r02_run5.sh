mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k rpart_kernel -s 4 -c 1 -o gpurun_out/r02_zipf_rpart python scratch/shard_zipf_prof.py 26 28 2 1 1.0 > gpurun_out/zipf_prof_ncu2.log 2>&1; echo "ncu2 rc=$?"
tail -3 gpurun_out/zipf_prof_ncu2.log
