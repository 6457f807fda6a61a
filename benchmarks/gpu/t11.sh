export CUDA_DEVICE_MAX_CONNECTIONS=32
python -m pytest tests/test_gpu_read.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -4
DNV_LANES=2 python benchmarks/ktrace.py 10
DNV_LANES=3 python benchmarks/ktrace.py 10
DNV_LANES=1 python benchmarks/ktrace.py 6
run() { # label, env..., args
  label=$1; shift
  env "$@" 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$label', 'step %.1f us' % (d['ms_per_step']*1e3), 'sync %.1f us' % (d['config']['ms_per_step_synchronous_api']*1e3), 'screen %.1f' % (d['roofline']['kernel_ms']*1e3), 'kernels %.1f' % (d['roofline']['step_kernels_ms']*1e3), d['config']['calls_per_step'])"
}
B="python bench.py --steps 40 --warmup 3 --contexts 1 --no-cpu-baseline"
for L in 1 2 3 4; do run L${L} DNV_LANES=$L $B; done
run L2_nograph DNV_GRAPHS=0 DNV_LANES=2 $B
run L3_nograph DNV_GRAPHS=0 DNV_LANES=3 $B
