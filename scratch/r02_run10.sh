mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/tpch_launches.csv python bench_ops.py --only tpch --sf 10 --steps 1 --warmup 1 > gpurun_out/tpch_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gb_shared_kernel -s 2 -c 1 -o gpurun_out/r02_q1 python bench_ops.py --only tpch --sf 10 --steps 1 --warmup 1 > gpurun_out/tpch_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 300 python -m pytest tests/test_gpu_fullsize.py::test_tpch_shapes_2_26_rows_device_resident -q -x 2>&1 | tail -4
