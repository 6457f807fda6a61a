python - <<'PY'
import sys, json
sys.path.insert(0, '.'); sys.path.insert(0, 'denovo-variant-caller-java_b200/python')
import numpy as np, torch
import denovo_b200 as D, bench
for x in bench.aux_undecided(D, torch, np, 0, [0.01, 0.05, 0.1, 0.25, 0.5, 1.0]): print(x)
PY
