#!/bin/bash
# A/B of round-2 kernel changes on one box: alternative builds (WVB_LIB) against the same synthetic batches
run() { lib=$1; shift; echo "== $lib: $*"; WVB_LIB=$PWD/wavpackdecoder_b200/$lib python tools/prof_run.py "$@" 2>&1 | grep -E "step [12]|flagged|Error|error"; }
for occ in 0 6; do
export WVB_FIXED_OCC=$occ
echo "WVB_FIXED_OCC=$occ"
for lib in libwvb.so libwvb_nostage.so; do
  run $lib --files 10000 --seconds 10 --steps 3
  run $lib --files 6000 --seconds 10 --steps 3
  run $lib --files 4700 --seconds 10 --steps 3
done
done
unset WVB_FIXED_OCC
for lib in libwvb.so libwvb_nostage.so; do
  run $lib --files 10000 --seconds 10 --steps 3 --kw terms=18,2,18,3,-2 deltas=2,2,2,2,2
  run $lib --files 10000 --seconds 10 --steps 3 --kw terms=18,17 deltas=2,2
done
