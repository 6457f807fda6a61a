DNV_PILEUP_S=8 timeout 120 python benchmarks/time_pileup.py > gpurun_out/p8.log 2>&1 && \
DNV_PILEUP_S=8 ncu --set full --clock-control none --import-source on -k regex:pileup_pipe -s 3 -c 1 -o gpurun_out/pp8 python benchmarks/time_pileup.py > gpurun_out/ncu8.log 2>&1
DNV_PILEUP_S=16 ncu --set full --clock-control none --import-source on -k regex:pileup_pipe -s 3 -c 1 -o gpurun_out/pp16 python benchmarks/time_pileup.py > gpurun_out/ncu16.log 2>&1
tail -2 gpurun_out/p8.log gpurun_out/ncu8.log gpurun_out/ncu16.log
