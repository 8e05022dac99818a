mkdir -p gpurun_out
timeout 300 python bench_ops.py --only tpch --sf 100 --steps 3 --warmup 2 > gpurun_out/r02_bench_ops_tpch_sf100.jsonl 2> gpurun_out/tpch.err; echo "tpch rc=$?"; tail -3 gpurun_out/tpch.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_ops_tpch_sf100.jsonl'):
    d = json.loads(l); print(d['op'][:20], d['rows'], round(d['ms'], 2), 'ms', round(d['rows_per_s'] / 1e9, 1), 'Grows/s frac', round(d['frac_of_peak'], 3))
PY
timeout 300 python -m pytest tests/test_gpu_operators.py -q -x 2>&1 | tail -3
