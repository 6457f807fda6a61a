python benchmarks/time_big_chunk.py
DNV_NO_LEVEL3=1 python benchmarks/time_big_chunk.py
DNV_LANES=1 python benchmarks/time_big_chunk.py
python benchmarks/time_big_chunk.py 10000000
