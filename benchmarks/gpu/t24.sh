for s in 8 16; do DNV_PILEUP_S=$s timeout 120 python benchmarks/time_pileup.py; done
for s in 8 16; do DNV_PILEUP_S=$s timeout 120 python benchmarks/time_pileup.py 8000000; done
V=$PWD/denovo-variant-caller-java_b200/lib/variants
for v in st3 st3k6 st4k6 st2k20; do for s in 8 16; do echo $v; DNV_B200_LIB=$V/$v.so DNV_PILEUP_S=$s timeout 120 python benchmarks/time_pileup.py; done; done
timeout 600 python -m pytest tests/test_gpu_pileup_mendel.py -m gpu -x -q 2>&1 | tail -3
