#!/bin/bash
# sort-kernel experiment: correctness, stage times per ranking variant, then ncu of the sort passes (run under gpurun)
mkdir -p gpurun_out
echo "== sort_check cluster 8"; timeout 60 python -u tools/sort_check.py 8 2>&1 | grep -v " ok " | tail -3
for cl in 0 1; do
  echo "== SORT_CLUSTER=8 SORT_CLAIM=$cl"
  SORT_CLAIM=$cl SORT_CLUSTER=8 timeout 100 python -u tools/time_stages.py 6000000 20 2>&1 | tail -2
done
SORT_CLUSTER=8 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sort_pass -s 7 -c 5 -o /tmp/sortcap_8 \
    python tools/profile_frame.py 6000000 4 > /dev/null 2>&1
ncu -i /tmp/sortcap_8.ncu-rep --page raw --csv > gpurun_out/sortcap_raw_8.csv 2>/dev/null
ncu -i /tmp/sortcap_8.ncu-rep --page source --csv --launch-skip 3 --launch-count 1 > gpurun_out/sortcap_src_8.csv 2>/dev/null
echo "== ncu cluster 8"; python tools/ncu_summary.py gpurun_out/sortcap_raw_8.csv
