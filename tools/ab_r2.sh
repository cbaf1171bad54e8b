#!/bin/bash
# A/B of round-2 kernel changes on one box: alternative builds (WVB_LIB) / settings against the same synthetic batches
run() { lib=$1; shift; echo "== $lib: $*"; WVB_LIB=$PWD/wavpackdecoder_b200/$lib python tools/prof_run.py "$@" 2>&1 | grep -E "step [12]|flagged|Error|error"; }
for lib in libwvb.so libwvb_cta64.so libwvb_cta32.so; do
  run $lib --files 10000 --seconds 10 --steps 3
  run $lib --files 6000 --seconds 10 --steps 3
  run $lib --files 6000 --seconds 10 --steps 3 --kw kind=2 bits=32
done
