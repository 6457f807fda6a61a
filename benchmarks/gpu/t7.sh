run() { # label, env..., args
  label=$1; shift
  env "$@" 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$label', 'step %.1f us' % (d['ms_per_step']*1e3), 'sync %.1f us' % (d['config']['ms_per_step_synchronous_api']*1e3), 'screen %.1f' % (d['roofline']['kernel_ms']*1e3), 'kernels %.1f' % (d['roofline']['step_kernels_ms']*1e3), d['config']['calls_per_step'])"
}
B="python bench.py --steps 40 --warmup 3 --contexts 1 --no-cpu-baseline"
for S in 2 3; do
run L3S${S}_notail DNV_DEBUG_SKIP=1 DNV_LANES=3 DNV_SCREEN_CTAS=$S $B
run L3S${S}_nofinish DNV_DEBUG_SKIP=2 DNV_LANES=3 DNV_SCREEN_CTAS=$S $B
run L2S${S}_notail DNV_DEBUG_SKIP=1 DNV_LANES=2 DNV_SCREEN_CTAS=$S $B
run L1S${S}_notail DNV_DEBUG_SKIP=1 DNV_LANES=1 DNV_SCREEN_CTAS=$S $B
run L1S${S}_nofinish DNV_DEBUG_SKIP=2 DNV_LANES=1 DNV_SCREEN_CTAS=$S $B
done
