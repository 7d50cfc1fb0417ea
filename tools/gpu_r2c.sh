#!/usr/bin/env bash
# Round-2 GPU pass C: whole GPU test-suite (incl. fused HPCF), c4 chain timing with the fused HPCF
set -u
o=gpurun_out
mkdir -p $o
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > $o/r2c_pytest.log 2>&1
tail -8 $o/r2c_pytest.log
ls -la $o | grep r2c
