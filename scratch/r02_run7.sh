mkdir -p gpurun_out
for r in 16 8; do
OLAP_GR_ROWS=$r timeout 200 python bench_ops.py --only groupby --groups 10000,100000,1000000,10000000,100000000 --steps 2 --warmup 1 > gpurun_out/gr_rows$r.jsonl 2> gpurun_out/gr_rows$r.err; echo "rows $r rc=$?"
python - gpurun_out/gr_rows$r.jsonl <<'PY'
import sys, json
for l in open(sys.argv[1]):
    d = json.loads(l); print(d['groups_hint'], round(d['ms'], 2), round(d['frac_of_peak'], 3), {k: round(v, 2) for k, v in d['phases_ms'].items() if v > 0.05})
PY
done
OLAP_GR_ROWS=8 timeout 300 python -m pytest tests/test_gpu_groupby_dense.py -q -x 2>&1 | tail -2
