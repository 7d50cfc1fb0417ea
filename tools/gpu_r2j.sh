#!/usr/bin/env bash
set -u
o=gpurun_out
mkdir -p $o
( time timeout 600 python -m pytest tests/test_gpu_partitioned.py -q -x ) > $o/r2j_pytest.log 2>&1
tail -4 $o/r2j_pytest.log
timeout 300 python bench.py --workload c3 --streams 4096 --steps 50 --warmup 20 --no-cpu-baseline --no-extras > $o/r2j_bench_c3.json 2> $o/r2j_bench_c3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file $o/r2j_launches_c3.csv \
    python tools/profile_run.py --workload c3 --streams 4096 --steps 4 > $o/r2j_ncu_list.log 2>&1
