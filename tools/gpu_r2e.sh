#!/usr/bin/env bash
# Round-2 GPU pass E: folded pre-renderer (c4 chain), whole GPU test-suite, c4 chain timings
set -u
o=gpurun_out
mkdir -p $o
( time timeout 900 python -m pytest tests -m gpu -q -x ) > $o/r2e_pytest.log 2>&1
tail -5 $o/r2e_pytest.log
for S in 1024 4096; do
  timeout 300 python tools/bench_c4_chain.py --streams $S --steps 100 > $o/r2e_c4_chain_s$S.json 2> $o/r2e_c4_chain_s$S.err
  timeout 300 python tools/bench_c4_chain.py --streams $S --steps 100 --fused > $o/r2e_c4_chain_fused_s$S.json 2> $o/r2e_c4_chain_fused_s$S.err
done
tail -c 400 $o/r2e_c4_chain_fused_s1024.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $o/r2e_launches_c4_fused.csv \
    python tools/bench_c4_chain.py --streams 1024 --steps 3 --warmup 2 --fused > $o/r2e_ncu_list.log 2>&1
ls -la $o | grep r2e
