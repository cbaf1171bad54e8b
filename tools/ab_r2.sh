#!/bin/bash
# A/B of round-2 kernel changes on one box: alternative builds (WVB_LIB) / settings against the same synthetic batches
run() { lib=$1; shift; echo "== $lib: $*"; WVB_LIB=$PWD/wavpackdecoder_b200/$lib python tools/prof_run.py "$@" 2>&1 | grep -E "step [12]|flagged|Error|error"; }
for lib in libwvb.so libwvb_hyb4.so libwvb_hyb5.so; do
  run $lib --files 8000 --seconds 10 --steps 3 --kw kind=1
  run $lib --files 12000 --seconds 10 --steps 3 --kw kind=1 channels=1 terms=18,18,2,3 deltas=2,2,2,2
done
