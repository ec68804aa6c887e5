#!/bin/bash
# K3 experiment probe (run under gpurun): stage times per env setting, then K3 under ncu with and without cache flushing
for o in 0 1; do echo "order $o"; B200GS_K3_ORDER=$o timeout 100 python -u tools/time_stages.py 6000000 24 2>&1 | tail -1; done
for cc in all none; do
echo "ncu cache-control $cc"
timeout 300 ncu --cache-control $cc --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,dram__bytes_read.sum \
  --clock-control none -k regex:k_composite -s 2 -c 1 python tools/profile_frame.py 6000000 4 2>&1 | grep -E "duration|issue_active|hit_rate|dram__bytes"
done
