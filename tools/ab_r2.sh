#!/bin/bash
# A/B of round-2 kernel changes on one box: alternative builds (WVB_LIB) / settings against the same synthetic batches
T16=18,18,2,3,-2,18,2,4,7,5,3,6,8,-1,18,2
D16=2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2
run() { lib=$1; shift; echo "== $lib: $*"; WVB_LIB=$PWD/wavpackdecoder_b200/$lib python tools/prof_run.py "$@" 2>&1 | grep -E "step [12]|flagged|Error|error"; }
for lib in libwvb.so; do
  run $lib --files 8000 --seconds 10 --steps 3 --kw kind=1
  run $lib --files 12000 --seconds 10 --steps 3 --kw kind=1 channels=1 terms=18,18,2,3 deltas=2,2,2,2
  run $lib --files 3000 --seconds 10 --steps 3 --open-flags 0x8 --kw bits=24 channels=6 sample_rate=48000 block_samples=24000 terms=$T16 deltas=$D16
done
