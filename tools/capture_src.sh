#!/bin/bash
# ncu full-set capture of the three heaviest kernels of one frame with source-level counters, exported as
# CSV (raw + source pages) under gpurun_out/.  Run under gpurun on one B200.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_preprocess|k_composite|k_bin' -s 6 -c 3 -o /tmp/src_full \
    python tools/profile_frame.py 6000000 4 > /dev/null 2>&1
ncu -i /tmp/src_full.ncu-rep --page raw --csv > gpurun_out/src_full_raw.csv 2>/dev/null
for k in k_preprocess k_composite k_bin; do
  ncu -i /tmp/src_full.ncu-rep --page source --csv -k regex:$k > gpurun_out/src_$k.csv 2>/dev/null
done
ls -la gpurun_out/src_*
