#!/bin/bash
# ncu evidence of one frame of the bench workload (6M Gaussians, 1920x1080), run under gpurun on one B200:
#   launches list (gpu__time_duration only, frames 2..5) and a full-set capture of one frame exported as CSV.
# usage: tools/capture_frame.sh <tag> <launches-per-frame>
TAG=${1:-r2}; L=${2:-9}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s $((2*L)) -c $((4*L)) --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/profile_frame.py 6000000 6 > /dev/null 2>&1
python tools/launch_table.py gpurun_out/${TAG}_launches.csv $L > gpurun_out/${TAG}_launches_table.txt
ncu --set full --clock-control none --import-source on -k regex:k_ -s $((2*L)) -c $L -o /tmp/${TAG}_frame_full \
    python tools/profile_frame.py 6000000 3 > /dev/null 2>&1
ncu -i /tmp/${TAG}_frame_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_frame_full_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_frame_full.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 > gpurun_out/${TAG}_src_k_preprocess.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_frame_full_raw.csv gpurun_out/${TAG}_frame_full_summary.json gpurun_out/${TAG}_traffic.json > gpurun_out/${TAG}_frame_full_summary.md
cat gpurun_out/${TAG}_launches_table.txt gpurun_out/${TAG}_frame_full_summary.md
