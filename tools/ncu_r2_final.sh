#!/bin/bash
# Round-2 final captures (one GPU).  Every ncu pass comes after the same command has exited 0 without ncu.
set -x
# 1. launch list of the bench command (kernel SHARE of the step; times under ncu are serialised and cold-cache)
python bench.py --steps 2 --warmup 1 > gpurun_out/r02_bench_plain.json 2> gpurun_out/r02_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_launches_bench.json 2> gpurun_out/r02_launches_bench.err
# 2. full capture of the headline kernel at bench size (tools/ncu_r2_extra.sh holds the other kernels' captures)
A="--files 10000 --seconds 10 --steps 2"
python tools/prof_run.py $A > gpurun_out/r02_fix_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_decode_pcm -s 1 -c 1 -o gpurun_out/r02_fix_bench python tools/prof_run.py $A > gpurun_out/r02_fix_ncu.log 2>&1
