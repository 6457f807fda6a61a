run() { # label, env..., args
  label=$1; shift
  env "$@" 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$label', 'step %.1f us' % (d['ms_per_step']*1e3), 'screen %.1f' % (d['roofline']['kernel_ms']*1e3))"
}
B="python bench.py --no-cpu-baseline --no-aux --warmup 3 --steps 20"
run K20 $B
run K20_nosampler BENCH_NO_SAMPLER=1 $B
run K20_nbuf1 BENCH_NBUF=1 $B
run K20_nbuf1_nosampler BENCH_NBUF=1 BENCH_NO_SAMPLER=1 $B
run K20_nbuf2 BENCH_NBUF=2 $B
run K20 $B
