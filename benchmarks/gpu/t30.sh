timeout 900 python -m pytest tests/test_gpu_pileup_mendel.py -m gpu -x -q -k "mendel" 2>&1 | tail -3
python - <<'PY'
import sys, os, json
sys.path.insert(0, '.')
sys.path.insert(0, 'denovo-variant-caller-java_b200/python')
import numpy as np, torch
import denovo_b200 as D
import bench
print(json.dumps(bench.aux_mendel(D, torch, np, 0), indent=1))
print(json.dumps(bench.aux_pileup(D, torch, np, 0), indent=1))
PY
