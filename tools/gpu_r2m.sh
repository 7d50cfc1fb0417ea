#!/usr/bin/env bash
# Round-2 GPU pass M: dual-signal render kernel for small batches (parity, then timing per batch size)
set -u
o=gpurun_out
mkdir -p $o
( time timeout 900 python -m pytest tests -m gpu -q -x ) > $o/r2m_pytest.log 2>&1
tail -4 $o/r2m_pytest.log
for S in 512 1024 1184 1536 2048; do
  for d in 0 1; do
    RTB_RENDER_DUAL=$d timeout 200 python bench.py --streams $S --steps 200 --warmup 20 --no-cpu-baseline --no-extras > $o/r2m_bench_s${S}_dual$d.json 2> $o/r2m_err.txt
    python -c "import json; d=json.load(open('$o/r2m_bench_s${S}_dual$d.json')); print('S=$S dual=$d ms', round(d['ms_per_step'],5), 'p99', round(d['p99_block_latency_ms'],5), 'render us', round(d['roofline']['us_per_launch'],2))"
  done
done
