#!/bin/bash
# A/B of round-2 kernel changes on one box: alternative builds (WVB_LIB) / settings against the same synthetic batches
run() { echo "== $*"; python tools/prof_run.py "$@" 2>&1 | grep -E "step [12]|flagged|Error|error"; }
for sp in 0 auto; do
  if [ $sp = auto ]; then unset WVB_SPREAD; else export WVB_SPREAD=$sp; fi
  echo "WVB_SPREAD=$sp"
  run --files 1 --seconds 60 --steps 3
  run --files 16 --seconds 10 --steps 3
  run --files 100 --seconds 10 --steps 3
  run --files 500 --seconds 10 --steps 3
  run --files 900 --seconds 10 --steps 3
  run --files 100 --seconds 10 --steps 3 --kw bits=24
  run --files 100 --seconds 10 --steps 3 --kw kind=1
done
