mkdir -p gpurun_out
( timeout 170 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py > gpurun_out/pytest_gpu_final2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final2.log )
tail -6 gpurun_out/pytest_gpu_final2.log
