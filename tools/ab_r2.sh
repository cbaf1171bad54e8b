#!/bin/bash
# A/B of round-2 kernel changes on one box: alternative builds (WVB_LIB) against the same synthetic batches
run() { lib=$1; shift; echo "== $lib: $*"; WVB_LIB=$PWD/wavpackdecoder_b200/$lib python tools/prof_run.py "$@" 2>&1 | grep -E "step [12]|flagged|Error|error"; }
for lib in libwvb_nostage.so libwvb.so; do
  run $lib --files 10000 --seconds 10 --steps 3
  run $lib --files 10000 --seconds 10 --steps 3 --kw terms=18,2,18,3,-2 deltas=2,2,2,2,2
  run $lib --files 6000 --seconds 10 --steps 3
  WVB_LIB=$PWD/wavpackdecoder_b200/$lib ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_write_lookup_miss.sum,smsp__inst_executed.sum --clock-control none -k regex:k_decode_pcm -s 1 -c 1 --csv --log-file gpurun_out/r2_dram_$lib.csv python tools/prof_run.py --files 10000 --seconds 10 --steps 2 > /dev/null 2>&1
  tail -7 gpurun_out/r2_dram_$lib.csv | cut -d, -f12-
done
