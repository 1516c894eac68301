timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in c4d_10 c4d_50; do python tools/profile_one.py $c --log2n 30 --reps 100 --warmup 2 --trend; sleep 2; done
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_c.json').read().strip().splitlines()[-1])
print(json.dumps(d['pipelines_summary']))
print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d.get('pipelines_error'))
PY
