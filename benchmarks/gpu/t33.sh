timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu_final.json 2> gpurun_out/r02_bench_1gpu_final.err; tail -c 600 gpurun_out/r02_bench_1gpu_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_1gpu_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['cpu_baseline'])
print(json.dumps(d.get('aux'),indent=0)[:3000])
PY
