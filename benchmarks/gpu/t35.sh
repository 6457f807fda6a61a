N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_${N}gpu_final.json 2> gpurun_out/r02_bench_${N}gpu_final.err
tail -c 300 gpurun_out/r02_bench_${N}gpu_final.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_${N}gpu_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['e2e']['value'], d['config'].get('gather_verified'), d['roofline']['whole_step_frac'])
a=d.get('aux') or {}
print(json.dumps(a.get('config4_genome_every_position')))
PY
