run() { # label, env..., args
  label=$1; shift
  env "$@" python bench.py --steps 40 --warmup 3 --contexts 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$label', 'step %.1f us' % (d['ms_per_step']*1e3), 'sync %.1f us' % (d['config']['ms_per_step_synchronous_api']*1e3), 'screen %.1f' % (d['roofline']['kernel_ms']*1e3), 'kernels %.1f' % (d['roofline']['step_kernels_ms']*1e3), 'e2e %.2f G' % (d['e2e']['value']/1e9), d['config']['calls_per_step'])"
}
python -m pytest tests/test_gpu_read.py -m gpu -x -q 2>&1 | tail -3
for L in 2 3 4; do for S in 2 3; do for B in 2 3; do run L${L}S${S}B${B} DNV_LANES=$L DNV_SCREEN_CTAS=$S DNV_SCREEN_BUFS=$B; done; done; done
run L1S3B2 DNV_LANES=1 DNV_SCREEN_CTAS=3 DNV_SCREEN_BUFS=2
run L1S2B3 DNV_LANES=1 DNV_SCREEN_CTAS=2 DNV_SCREEN_BUFS=3
run L1S4B2 DNV_LANES=1 DNV_SCREEN_CTAS=4 DNV_SCREEN_BUFS=2
