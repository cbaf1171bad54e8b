#!/bin/bash
# Round-2 final captures (one GPU).  Every ncu pass comes after the same command has exited 0 without ncu.
set -x
T16=18,18,2,3,-2,18,2,4,7,5,3,6,8,-1,18,2
D16=2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2
# 1. launch list of the bench command (kernel SHARE of the step; times under ncu are serialised and cold-cache)
python bench.py --steps 2 --warmup 1 > gpurun_out/r02_bench_plain.json 2> gpurun_out/r02_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_launches_bench.json 2> gpurun_out/r02_launches_bench.err
# 2. full captures
A="--files 10000 --seconds 10 --steps 2"
python tools/prof_run.py $A > gpurun_out/r02_fix_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_decode_pcm -s 1 -c 1 -o gpurun_out/r02_fix_bench python tools/prof_run.py $A > gpurun_out/r02_fix_ncu.log 2>&1
B="--files 1000 --seconds 10 --steps 2 --kw kind=3 dsd_mode=3 block_samples=22050"
python tools/prof_run.py $B > gpurun_out/r02_dsdhigh_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_dsd_high -s 1 -c 1 -o gpurun_out/r02_dsdhigh python tools/prof_run.py $B > gpurun_out/r02_dsdhigh_ncu.log 2>&1
C="--files 1000 --seconds 10 --steps 2 --kw kind=3 dsd_mode=1 block_samples=22050"
python tools/prof_run.py $C > gpurun_out/r02_dsdfast_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_dsd_fast_dec -s 1 -c 1 -o gpurun_out/r02_dsdfast python tools/prof_run.py $C > gpurun_out/r02_dsdfast_ncu.log 2>&1
D="--files 2000 --seconds 10 --steps 2 --open-flags 0x10000 --kw bits=24 channels=6 sample_rate=48000 block_samples=24000 terms=$T16 deltas=$D16"
python tools/prof_run.py $D > gpurun_out/r02_t16_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_decode_pcm -s 1 -c 1 -o gpurun_out/r02_t16 python tools/prof_run.py $D > gpurun_out/r02_t16_ncu.log 2>&1
