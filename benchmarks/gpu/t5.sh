run() { # label, env..., args
  label=$1; shift
  env "$@" python bench.py --steps 40 --warmup 3 --contexts 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$label', 'step %.1f us' % (d['ms_per_step']*1e3), 'sync %.1f us' % (d['config']['ms_per_step_synchronous_api']*1e3), 'screen %.1f' % (d['roofline']['kernel_ms']*1e3), 'kernels %.1f' % (d['roofline']['step_kernels_ms']*1e3), 'e2e %.2f G' % (d['e2e']['value']/1e9), d['config']['calls_per_step'])"
}
for B in 2 3 4; do run L1S1B${B} DNV_LANES=1 DNV_SCREEN_CTAS=1 DNV_SCREEN_BUFS=$B; done
for L in 2 3; do for B in 3 4; do run L${L}S1B${B} DNV_LANES=$L DNV_SCREEN_CTAS=1 DNV_SCREEN_BUFS=$B; done; done
run L3S2B2 DNV_LANES=3 DNV_SCREEN_CTAS=2 DNV_SCREEN_BUFS=2
