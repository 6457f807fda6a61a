timeout 2400 python -m pytest tests/test_gpu_read.py tests/test_gpu_scale.py tests/test_gpu_pileup_mendel.py -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import sys, json
sys.path.insert(0, '.'); sys.path.insert(0, 'denovo-variant-caller-java_b200/python')
import numpy as np, torch
import denovo_b200 as D, bench
for m, lrt in (("MAP", 1.0), ("LRT", 10.0), ("BAYES", 1.0)):
    print(json.dumps(bench.aux_read_config(D, torch, 0, m, 30, 1e-2, 1e-8, lrt, 10_000_000, 0x5EED0003)))
PY
