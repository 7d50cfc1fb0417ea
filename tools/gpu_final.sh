#!/usr/bin/env bash
# What the driver runs at round end, on one GPU: GPU tests, smoke(), both bench arms.
set -u
o=gpurun_out
mkdir -p $o
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > $o/final_pytest.log 2>&1
tail -4 $o/final_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $o/final_smoke.log 2>&1
tail -3 $o/final_smoke.log
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $o/final_bench_reference.json 2> $o/final_bench_reference.err
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $o/final_bench.json 2> $o/final_bench.err
tail -3 $o/final_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/final_launches_bench_c2.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $o/final_ncu_list.log 2>&1
ls -la $o | grep final
