python - <<'PY'
import sys, json
sys.path.insert(0, '.'); sys.path.insert(0, 'denovo-variant-caller-java_b200/python')
import numpy as np, torch
import denovo_b200 as D, bench
stream = torch.cuda.Stream()
with D.TrioEngine(1e-2, 1e-8, 1.0, "BAYES") as eng:
    eng.set_stream(stream.cuda_stream)
    for rep in range(2):
        g = bench.genome_walk(D, torch, np, None, eng, stream, None, 0, 1, 0)
        print(rep, g)
PY
