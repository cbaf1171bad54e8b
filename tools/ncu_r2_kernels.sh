set -x
T16=18,18,2,3,-2,18,2,4,7,5,3,6,8,-1,18,2
D16=2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2
A="--files 1000 --seconds 10 --steps 2 --kw kind=3 dsd_mode=1 block_samples=22050"
python tools/prof_run.py $A > gpurun_out/r2_dsdfast_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_dsd_fast_dec -s 1 -c 1 -o gpurun_out/r2_dsdfast python tools/prof_run.py $A > gpurun_out/r2_dsdfast_ncu.log 2>&1
B="--files 1000 --seconds 10 --steps 2 --kw kind=3 dsd_mode=3 block_samples=22050"
python tools/prof_run.py $B > gpurun_out/r2_dsdhigh_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_dsd_high -s 1 -c 1 -o gpurun_out/r2_dsdhigh python tools/prof_run.py $B > gpurun_out/r2_dsdhigh_ncu.log 2>&1
C="--files 3000 --seconds 10 --steps 2 --open-flags 0x8 --kw bits=24 channels=6 sample_rate=48000 block_samples=24000 terms=$T16 deltas=$D16"
python tools/prof_run.py $C > gpurun_out/r2_t16_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_decode_pcm -s 1 -c 1 -o gpurun_out/r2_t16 python tools/prof_run.py $C > gpurun_out/r2_t16_ncu.log 2>&1
