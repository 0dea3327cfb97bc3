#!/bin/bash
# Round-2 ncu captures (run on the GPU box: gpurun -- bash tools/profile_r02.sh).  Numbers printed by runs under ncu are
# never bench values; the summaries are produced afterwards with tools/ncu_summarize.py and committed under profiles/.
set -u
mkdir -p gpurun_out
# 1. launch list of the bench command itself: per-launch time and DRAM bytes of every kernel in the step (10 GiB workload)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv \
    --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/launches_r02.stdout 2>&1
# 2. full set of one launch of every kernel family (2 GiB per launch keeps the replay short)
ncu --set full --clock-control none --import-source on -o gpurun_out/prof_r02 -f \
    -k regex:'encode_kernel|decode_kernel|hamming_kernel|complement_kernel|reverse_complement|b5_encode_kernel|b5_decode_kernel' \
    python tools/profile_targets.py > gpurun_out/prof_r02.stdout 2>&1
ncu -i gpurun_out/prof_r02.ncu-rep --page raw --csv > gpurun_out/prof_r02_raw.csv 2> gpurun_out/prof_r02_raw.err
ls -la gpurun_out/prof_r02* | head
