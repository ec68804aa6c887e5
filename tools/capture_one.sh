#!/bin/bash
# ncu full-set capture of ONE kernel of a frame with source-level counters: tools/capture_one.sh <kernel-regex> [skip]
# Outputs gpurun_out/one_raw.csv and gpurun_out/one_src.csv.  Run under gpurun on one B200.
mkdir -p gpurun_out
ncu -f --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-2} -c 1 -o /tmp/one_full \
    python tools/profile_frame.py 6000000 4 > /dev/null 2>&1
ncu -i /tmp/one_full.ncu-rep --page raw --csv > gpurun_out/one_raw.csv 2>/dev/null
ncu -i /tmp/one_full.ncu-rep --page source --csv > gpurun_out/one_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/one_raw.csv
