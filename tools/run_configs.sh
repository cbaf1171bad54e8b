#!/bin/bash
# Device-resident decode throughput for the other BASELINE.json configs (3: 24-bit 5.1 16 terms; 4: float / int32 / hybrid; 5: DSD64).
# Prints one line per config; used to fill DESIGN.md section 6b.  Run on the GPU box: bash tools/run_configs.sh
set -u
T16=18,18,2,3,-2,18,2,4,7,5,3,6,8,-1,18,2
D16=2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2
run() { echo "== $1"; shift; python tools/prof_run.py "$@" 2>&1 | grep -E "step 1|flagged|Error|error" ; }
run "config3 24-bit 48k 5.1, 16 terms, OPEN_2CH_MAX (reference parity: FL/FR only)" --files 6000 --seconds 10 --open-flags 0x8 --kw bits=24 channels=6 sample_rate=48000 block_samples=24000 terms=$T16 deltas=$D16
run "config3 24-bit 48k 5.1, 16 terms, all six channels (extension)" --files 2000 --seconds 10 --open-flags 0x10000 --kw bits=24 channels=6 sample_rate=48000 block_samples=24000 terms=$T16 deltas=$D16
run "config4a float-flagged (24-bit mantissas, FloatUtils shift/clip)" --files 6000 --seconds 10 --kw kind=2 bits=32
run "config4b int32 + WVX (8 sent bits)" --files 6000 --seconds 10 --kw bits=32 int32_sent_bits=8
run "config4c hybrid 4 bits/sample stereo" --files 8000 --seconds 10 --kw kind=1
run "config4c hybrid mono" --files 12000 --seconds 10 --kw kind=1 channels=1 terms=18,18,2,3 deltas=2,2,2,2
run "config5 DSD64 stereo mode 0 (raw)" --files 1000 --seconds 10 --kw kind=3 dsd_mode=0 block_samples=22050
run "config5 DSD64 stereo mode 1 (fast)" --files 1000 --seconds 10 --kw kind=3 dsd_mode=1 block_samples=22050
run "config5 DSD64 stereo mode 3 (high)" --files 1000 --seconds 10 --kw kind=3 dsd_mode=3 block_samples=22050
run "16-bit stereo, generic kernel (terms 18,18,2,3,-2 with a different order)" --files 10000 --seconds 10 --kw terms=18,2,18,3,-2 deltas=2,2,2,2,2
run "16-bit stereo, FFmpeg default-level list 18,18,2,17,3 (in-register kernel V_FIXED_B)" --files 10000 --seconds 10 --kw terms=18,18,2,17,3 deltas=2,2,2,2,2
run "16-bit stereo, FFmpeg fastest-level list 18,17 (in-register kernel V_FIXED_C)" --files 10000 --seconds 10 --kw terms=18,17 deltas=2,2
run "24-bit stereo 44.1k, stock terms (shift-free, plain kernel)" --files 8000 --seconds 10 --kw bits=24
run "16-bit mono, stock mono terms (in-register kernel)" --files 16000 --seconds 10 --kw channels=1 terms=18,18,2,3 deltas=2,2,2,2
