#!/usr/bin/env bash
# Round-2 GPU pass B: the tensor-core partition MAC (parity first, then c3 timing and an ncu capture).
set -u
o=gpurun_out
mkdir -p $o
( time timeout 600 python -m pytest tests/test_gpu_partitioned.py tests/test_gpu_dropin.py -x -q -s ) > $o/r2b_pytest.log 2>&1
tail -15 $o/r2b_pytest.log
for tc in 1 0; do
  RTB_PART_TC=$tc timeout 300 python bench.py --workload c3 --streams 4096 --steps 50 --warmup 20 --no-cpu-baseline --no-extras \
     > $o/r2b_bench_c3_tc$tc.json 2> $o/r2b_bench_c3_tc$tc.err
  tail -c 300 $o/r2b_bench_c3_tc$tc.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $o/r2b_launches_c3.csv \
    python tools/profile_run.py --workload c3 --streams 4096 --steps 5 > $o/r2b_ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sh_partition_tc -s 2 -c 1 -f -o $o/prof_r2b_pt \
    python tools/profile_run.py --workload c3 --streams 4096 --steps 4 > $o/r2b_ncu_pt.log 2>&1
ls -la $o | grep r2b
