#!/usr/bin/env bash
# Round-2 GPU pass G: m-grouped render kernel for 32 / 64-sample blocks (tests, c4 timings)
set -u
o=gpurun_out
mkdir -p $o
( time timeout 900 python -m pytest tests -m gpu -q -x ) > $o/r2g_pytest.log 2>&1
tail -5 $o/r2g_pytest.log
timeout 300 python bench.py --workload c4 --streams 1024 --steps 200 --warmup 20 --no-cpu-baseline --no-extras > $o/r2g_bench_c4.json 2> $o/r2g_bench_c4.err
RTB_RENDER_MG=0 timeout 300 python bench.py --workload c4 --streams 1024 --steps 200 --warmup 20 --no-cpu-baseline --no-extras > $o/r2g_bench_c4_termwise.json 2>> $o/r2g_bench_c4.err
for S in 1024 4096; do
  timeout 300 python tools/bench_c4_chain.py --streams $S --steps 100 --fused > $o/r2g_c4_chain_fused_s$S.json 2> $o/r2g_c4_chain_fused_s$S.err
done
ls -la $o | grep r2g
