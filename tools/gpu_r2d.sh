#!/usr/bin/env bash
# Round-2 GPU pass D: restructured tensor-core partition MAC (parity, c3 timing, ncu), device SH transform test
set -u
o=gpurun_out
mkdir -p $o
( time timeout 600 python -m pytest tests/test_gpu_partitioned.py tests/test_gpu_api.py tests/test_gpu_hpcf_fused.py -x -q -s ) > $o/r2d_pytest.log 2>&1
tail -6 $o/r2d_pytest.log
timeout 300 python bench.py --workload c3 --streams 4096 --steps 50 --warmup 20 --no-cpu-baseline --no-extras \
     > $o/r2d_bench_c3.json 2> $o/r2d_bench_c3.err
tail -c 300 $o/r2d_bench_c3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $o/r2d_launches_c3.csv \
    python tools/profile_run.py --workload c3 --streams 4096 --steps 5 > $o/r2d_ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"sh_partition_tc|sh_spectra_soa" -s 4 -c 2 -f -o $o/prof_r2d_pt \
    python tools/profile_run.py --workload c3 --streams 4096 --steps 4 > $o/r2d_ncu_pt.log 2>&1
ls -la $o | grep r2d
