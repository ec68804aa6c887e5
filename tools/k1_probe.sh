#!/bin/bash
# K1 experiment probe (run under gpurun): stage times, then instruction count / duration / issue utilisation of one K1 launch
mkdir -p gpurun_out
timeout 200 python -u tools/time_stages.py 6000000 24 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
  --clock-control none -k regex:k_preprocess -s 2 -c 2 python tools/profile_frame.py 6000000 4 2>&1 | grep -E "k_preprocess|duration|inst_executed|issue_active|registers"
