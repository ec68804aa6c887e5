#!/bin/bash
# Round-1 ncu evidence (run under gpurun on one B200): per-launch device times of whole frames, a full-set
# capture of the 10 kernels of one frame exported to CSV on the box (the .ncu-rep is too big to travel
# back), and compute-sanitizer memcheck / racecheck of a small multi-model frame.  Outputs under gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
L=10   # kernel launches per frame: preprocess, 4 depth-sort passes, bin, tile finish, 2 tile-sort passes, composite
# 6M @1080p, views 0..5 of the bench batch; frames 2..5
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s $((2*L)) -c $((4*L)) --csv --log-file gpurun_out/r1_launches.csv \
    python tools/profile_frame.py 6000000 6 > /dev/null 2>&1
python tools/launch_table.py gpurun_out/r1_launches.csv $L > gpurun_out/r1_launches_table.txt
ncu --set full --clock-control none --import-source on -k regex:k_ -s $((2*L)) -c $L -o /tmp/r1_frame_full \
    python tools/profile_frame.py 6000000 3 > /dev/null 2>&1
ncu -i /tmp/r1_frame_full.ncu-rep --page raw --csv > gpurun_out/r1_frame_full_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r1_frame_full_raw.csv gpurun_out/r1_frame_full_summary.json > gpurun_out/r1_frame_full_summary.md
( echo '$ compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_frame.py'
  compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_frame.py 2>&1 | tail -3
  echo; echo '$ compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_frame.py'
  compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_frame.py 2>&1 | tail -3 ) > gpurun_out/r1_sanitizer_raw.txt
cat gpurun_out/r1_launches_table.txt gpurun_out/r1_frame_full_summary.md gpurun_out/r1_sanitizer_raw.txt
