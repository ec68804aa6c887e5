#!/bin/bash
# Round-1 ncu evidence (run under gpurun on one B200): per-launch device times of whole frames, and a
# full-set capture of the 12 kernels of one frame, exported to CSV on the box (the .ncu-rep is too big
# to travel back).  Outputs under gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
# 6M @1080p, views 0..5 of the bench batch; 12 kernel launches per frame
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 24 -c 48 --csv --log-file gpurun_out/r1_launches.csv \
    python tools/profile_frame.py 6000000 6 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 24 -c 12 -o /tmp/r1_frame_full \
    python tools/profile_frame.py 6000000 3 > /dev/null 2>&1
ncu -i /tmp/r1_frame_full.ncu-rep --page raw --csv > gpurun_out/r1_frame_full_raw.csv 2>/dev/null
ncu -i /tmp/r1_frame_full.ncu-rep --page details --csv > gpurun_out/r1_frame_full_details.csv 2>/dev/null
ls -la gpurun_out /tmp/r1_frame_full.ncu-rep
