#!/usr/bin/env bash
# Round-2 GPU pass H: whole GPU test-suite with the regression bars, MN-major TF32 probe (retry), fd_render ncu
set -u
o=gpurun_out
mkdir -p $o
( time timeout 900 python -m pytest tests -m gpu -q -x ) > $o/r2h_pytest.log 2>&1
tail -5 $o/r2h_pytest.log
timeout 60 ./build/tc_probe_mn > $o/r2h_tc_probe_mn.txt 2>&1
cat $o/r2h_tc_probe_mn.txt
timeout 120 python tools/profile_fd.py --steps 50 > $o/r2h_fd_render.json 2> $o/r2h_fd_render.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fd_render -s 10 -c 1 -f -o $o/prof_r2h_fd \
    python tools/profile_fd.py --steps 3 > $o/r2h_ncu_fd.log 2>&1
ls -la $o | grep r2h
