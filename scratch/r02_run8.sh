mkdir -p gpurun_out
for r in 16 8; do
OLAP_RPART_ROWS=$r timeout 200 python bench.py --steps 6 --warmup 3 --skip-e2e --skip-cpu --skip-general --skip-groupby --skip-payload > gpurun_out/rpart_rows$r.json 2> gpurun_out/rpart_rows$r.err; echo "rows $r rc=$?"
python - gpurun_out/rpart_rows$r.json <<'PY'
import sys, json
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['phases_ms'].items()})
PY
done
OLAP_RPART_ROWS=8 timeout 300 python -m pytest tests/test_gpu_djoin.py -q -x 2>&1 | tail -2
timeout 200 python bench_ops.py --only groupby --groups 10000,100000,1000000,10000000,100000000 --steps 2 --warmup 1 > gpurun_out/gr_auto.jsonl 2> gpurun_out/gr_auto.err; echo "gb rc=$?"
python - gpurun_out/gr_auto.jsonl <<'PY'
import sys, json
for l in open(sys.argv[1]):
    d = json.loads(l); print(d['groups_hint'], round(d['ms'], 2), round(d['frac_of_peak'], 3), {k: round(v, 2) for k, v in d['phases_ms'].items() if v > 0.05})
PY
