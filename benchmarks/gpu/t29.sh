echo skip tests
python - <<'PY'
import sys, os, json
sys.path.insert(0, '.')
sys.path.insert(0, 'denovo-variant-caller-java_b200/python')
import numpy as np, torch
import denovo_b200 as D
import bench
r = bench.aux_pileup(D, torch, np, 0)
print(json.dumps(r, indent=1))
PY
