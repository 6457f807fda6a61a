python -m pytest tests/test_gpu_read.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -4
export CUDA_DEVICE_MAX_CONNECTIONS=32
DNV_LANES=1 python benchmarks/ktrace.py 5 | tail -8
python benchmarks/ktrace.py 16 | tail -20
run() { label=$1; shift; env "$@" 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$label', 'step %.1f us' % (d['ms_per_step']*1e3), 'sync %.1f' % (d['config']['ms_per_step_synchronous_api']*1e3), 'screen %.1f' % (d['roofline']['kernel_ms']*1e3), 'kernels %.1f' % (d['roofline']['step_kernels_ms']*1e3), 'undecided', d['roofline']['sites_to_fp64_kernel'], 'calls', d['config']['calls_per_step'], 'whole-step frac %.3f' % d['roofline']['whole_step_frac'])"; }
B="python bench.py --no-cpu-baseline --no-aux --warmup 3"
run K20 $B --steps 20
run K100 $B --steps 100
run K20_L4 DNV_LANES=4 $B --steps 20
run K20_L2 DNV_LANES=2 $B --steps 20
