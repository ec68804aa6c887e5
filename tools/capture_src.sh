#!/bin/bash
# ncu full-set capture of selected kernels of one frame with source-level (SASS) counters, exported as CSV under
# gpurun_out/: tools/capture_src.sh "<kernel-regex>" <count> [skip].  Run under gpurun on one B200.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-0} -c $2 -o /tmp/src_full \
    python tools/profile_frame.py 6000000 4 > /dev/null 2>&1
ncu -i /tmp/src_full.ncu-rep --page raw --csv > gpurun_out/src_full_raw.csv 2>/dev/null
for i in $(seq 0 $(($2-1))); do
  ncu -i /tmp/src_full.ncu-rep --page source --csv --launch-skip $i --launch-count 1 > gpurun_out/src_$i.csv 2>/dev/null
done
python tools/ncu_summary.py gpurun_out/src_full_raw.csv
