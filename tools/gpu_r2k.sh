#!/usr/bin/env bash
set -u
o=gpurun_out
mkdir -p $o
for d in 0 1 2 4 3 7; do
  RTB_PT_DEBUG=$d timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sh_partition_tc -c 3 --csv --log-file $o/r2k_dbg$d.csv \
    python tools/profile_run.py --workload c3 --streams 4096 --steps 4 > $o/r2k_dbg$d.log 2>&1
  echo "debug=$d $(grep sh_partition_tc $o/r2k_dbg$d.csv | awk -F'\",\"' '{print $NF}' | tr -d '\"' | tr '\n' ' ')"
done
