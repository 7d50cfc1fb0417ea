#!/usr/bin/env bash
set -u
o=gpurun_out
mkdir -p $o
( time timeout 600 python -m pytest tests/test_gpu_partitioned.py -q -x -k "c3 or tc" ) > $o/r2l_pytest.log 2>&1
tail -4 $o/r2l_pytest.log
for d in 0 1 2 7; do
  RTB_PT_DEBUG=$d timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sh_partition_tc -c 3 --csv --log-file $o/r2l_dbg$d.csv \
    python tools/profile_run.py --workload c3 --streams 4096 --steps 4 > $o/r2l_dbg$d.log 2>&1
  echo "debug=$d $(grep sh_partition_tc $o/r2l_dbg$d.csv | awk -F'\",\"' '{print $NF}' | tr -d '\"' | tr '\n' ' ')"
done
timeout 300 python bench.py --workload c3 --streams 4096 --steps 50 --warmup 20 --no-cpu-baseline --no-extras > $o/r2l_bench_c3.json 2> $o/r2l_bench_c3.err
python -c "import json; d=json.load(open('$o/r2l_bench_c3.json')); print('c3', d['value'], d['ms_per_step'])"
