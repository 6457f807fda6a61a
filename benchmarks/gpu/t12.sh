export CUDA_DEVICE_MAX_CONNECTIONS=32
python -m pytest tests/test_gpu_read.py -m gpu -x -q 2>&1 | tail -2
run() { # label, env..., args
  label=$1; shift
  env "$@" 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$label', 'step %.1f us' % (d['ms_per_step']*1e3), 'screen %.1f' % (d['roofline']['kernel_ms']*1e3), 'kernels %.1f' % (d['roofline']['step_kernels_ms']*1e3), d['config']['calls_per_step'])"
}
B="python bench.py --steps 40 --warmup 3 --contexts 1 --no-cpu-baseline"
for L in 2 3 4; do run L${L}_p0 DNV_LANES=$L $B; run L${L}_p1 DNV_TAIL_PRIO=1 DNV_LANES=$L $B; done
run L3_p0_conn8 CUDA_DEVICE_MAX_CONNECTIONS=8 DNV_LANES=3 $B
run L3_p0_S3 DNV_SCREEN_CTAS=3 DNV_LANES=3 $B
DNV_LANES=3 python benchmarks/ktrace.py 10
