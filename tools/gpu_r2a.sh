#!/usr/bin/env bash
# Round-2 GPU pass A: GPU tests, driver-shaped bench lines (both arms), ncu captures of the FDL,
# partition-MAC and direction-switching kernels.  Usage on the GPU box: bash tools/gpu_r2a.sh
set -u
o=gpurun_out
mkdir -p $o
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $o/r2a_pytest.log 2>&1
tail -5 $o/r2a_pytest.log
( time timeout 400 python bench.py --steps 20 --warmup 5 ) > $o/r2a_bench.json 2> $o/r2a_bench.err
tail -c 600 $o/r2a_bench.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > $o/r2a_bench_reference.json 2> $o/r2a_bench_reference.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ols_fdl -s 24 -c 1 -f -o $o/prof_r2a_fdl \
    python tools/bench_fdl.py --streams 16384 --steps 4 --warmup 2 > $o/r2a_ncu_fdl.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:sh_partition_mac -s 2 -c 1 -f -o $o/prof_r2a_mac \
    python tools/profile_run.py --workload c3 --streams 4096 --steps 4 > $o/r2a_ncu_mac.log 2>&1
ls -la $o | grep r2a
