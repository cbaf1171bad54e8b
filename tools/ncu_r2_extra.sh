T16=18,18,2,3,-2,18,2,4,7,5,3,6,8,-1,18,2
D16=2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2
D="--files 3000 --seconds 10 --steps 2 --open-flags 0x8 --kw bits=24 channels=6 sample_rate=48000 block_samples=24000 terms=$T16 deltas=$D16"
python tools/prof_run.py $D > gpurun_out/r02_t16fix_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_decode_pcm -s 1 -c 1 -o gpurun_out/r02_t16fix python tools/prof_run.py $D > gpurun_out/r02_t16fix_ncu.log 2>&1
E="--files 6000 --seconds 10 --steps 2 --kw kind=2 bits=32"
python tools/prof_run.py $E > gpurun_out/r02_floatfix_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_decode_pcm -s 1 -c 1 -o gpurun_out/r02_floatfix python tools/prof_run.py $E > gpurun_out/r02_floatfix_ncu.log 2>&1
