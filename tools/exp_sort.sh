#!/bin/bash
# sort-kernel experiment: stage times for every cluster size / ranking variant, then ncu of the sort passes (run under gpurun)
mkdir -p gpurun_out
timeout 60 python -u tools/sort_check.py 8 2>&1 | tail -2
for cl in 0 1; do for c in 8 4 2 1; do
  echo "== SORT_CLUSTER=$c SORT_CLAIM=$cl"
  SORT_CLAIM=$cl SORT_CLUSTER=$c timeout 100 python -u tools/time_stages.py 6000000 20 2>&1 | tail -2
done; done
for c in 8 1; do
  SORT_CLUSTER=$c timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sort_pass -s 7 -c 5 -o /tmp/sortcap_$c \
      python tools/profile_frame.py 6000000 4 > /dev/null 2>&1
  ncu -i /tmp/sortcap_$c.ncu-rep --page raw --csv > gpurun_out/sortcap_raw_$c.csv 2>/dev/null
  ncu -i /tmp/sortcap_$c.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 > gpurun_out/sortcap_src_$c.csv 2>/dev/null
  echo "== ncu cluster $c"; python tools/ncu_summary.py gpurun_out/sortcap_raw_$c.csv
done
