run() { # label, env..., args
  label=$1; shift
  env "$@" 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$label', 'step %.1f us' % (d['ms_per_step']*1e3), 'sync %.1f us' % (d['config']['ms_per_step_synchronous_api']*1e3), 'screen %.1f' % (d['roofline']['kernel_ms']*1e3), 'kernels %.1f' % (d['roofline']['step_kernels_ms']*1e3), 'e2e %.2f G' % (d['e2e']['value']/1e9), d['config']['calls_per_step'])"
}
B="python bench.py --steps 40 --warmup 3 --contexts 1 --no-cpu-baseline"
run L4S2_conn32 CUDA_DEVICE_MAX_CONNECTIONS=32 DNV_LANES=4 DNV_SCREEN_CTAS=2 $B
run L4S3_conn32 CUDA_DEVICE_MAX_CONNECTIONS=32 DNV_LANES=4 DNV_SCREEN_CTAS=3 $B
run L3S2_conn32 CUDA_DEVICE_MAX_CONNECTIONS=32 DNV_LANES=3 DNV_SCREEN_CTAS=2 $B
run L3S3_conn32 CUDA_DEVICE_MAX_CONNECTIONS=32 DNV_LANES=3 DNV_SCREEN_CTAS=3 $B
run L2S2_conn32 CUDA_DEVICE_MAX_CONNECTIONS=32 DNV_LANES=2 DNV_SCREEN_CTAS=2 $B
run L1S3_40M DNV_LANES=1 DNV_SCREEN_CTAS=3 $B --sites 40000000
run L1S2_40M DNV_LANES=1 DNV_SCREEN_CTAS=2 $B --sites 40000000
run L1S4_40M DNV_LANES=1 DNV_SCREEN_CTAS=4 $B --sites 40000000
run L3S2_40M DNV_LANES=3 DNV_SCREEN_CTAS=2 $B --sites 40000000
