python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_t3_pytest.log; tail -4 gpurun_out/r02_t3_pytest.log
run() { # label, env..., args
  label=$1; shift
  env "$@" python bench.py --steps 40 --warmup 3 --contexts 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$label', 'step %.1f us' % (d['ms_per_step']*1e3), 'sync %.1f us' % (d['config']['ms_per_step_synchronous_api']*1e3), 'screen %.1f' % (d['roofline']['kernel_ms']*1e3), 'kernels %.1f' % (d['roofline']['step_kernels_ms']*1e3), 'e2e %.2f G' % (d['e2e']['value']/1e9), d['config']['calls_per_step'])"
}
run L4S3 DNV_LANES=4 DNV_SCREEN_CTAS=3
run L2S3 DNV_LANES=2 DNV_SCREEN_CTAS=3
run L3S3 DNV_LANES=3 DNV_SCREEN_CTAS=3
run L1S3 DNV_LANES=1 DNV_SCREEN_CTAS=3
run L2S4 DNV_LANES=2 DNV_SCREEN_CTAS=4
run L4S4 DNV_LANES=4 DNV_SCREEN_CTAS=4
run L1S4 DNV_LANES=1 DNV_SCREEN_CTAS=4
