run() { # label, env..., args
  label=$1; shift
  env "$@" 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$label', 'step %.1f us' % (d['ms_per_step']*1e3), 'sync %.1f' % (d['config']['ms_per_step_synchronous_api']*1e3), 'screen %.1f' % (d['roofline']['kernel_ms']*1e3), 'kernels %.1f' % (d['roofline']['step_kernels_ms']*1e3), d['config']['calls_per_step'])"
}
B="python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-aux"
run base $B
run base_again $B
for G in 1 2 3 4; do run tailgrid$G DNV_TAIL_GRID=$G $B; done
run lanes2 DNV_LANES=2 $B
run lanes4 DNV_LANES=4 $B
run steps20 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-aux
run steps100 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-aux
