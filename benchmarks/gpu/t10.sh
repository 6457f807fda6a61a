export CUDA_DEVICE_MAX_CONNECTIONS=32
DNV_LANES=3 DNV_SCREEN_CTAS=2 python benchmarks/ktrace.py 10
DNV_LANES=3 DNV_SCREEN_CTAS=3 python benchmarks/ktrace.py 10
DNV_LANES=2 DNV_SCREEN_CTAS=2 python benchmarks/ktrace.py 10
DNV_GRAPHS=0 DNV_LANES=3 DNV_SCREEN_CTAS=2 python benchmarks/ktrace.py 10
