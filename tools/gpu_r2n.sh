#!/usr/bin/env bash
set -u
o=gpurun_out
mkdir -p $o
( time timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_gpu_sh_chain.py tests/test_gpu_hpcf_fused.py -q -x ) > $o/r2n_pytest.log 2>&1
tail -3 $o/r2n_pytest.log
for S in 512 1024 2048; do
  for mg in default 1; do
    if [ $mg = default ]; then unset RTB_RENDER_MG; else export RTB_RENDER_MG=1; fi
    timeout 200 python bench.py --workload c4 --streams $S --steps 200 --warmup 20 --no-cpu-baseline --no-extras > $o/r2n_c4_s${S}_mg$mg.json 2> $o/r2n_err.txt
    python -c "import json; d=json.load(open('$o/r2n_c4_s${S}_mg$mg.json')); print('c4 S=$S mg=$mg ms', round(d['ms_per_step'],5), d['roofline']['kernel'], round(d['roofline']['us_per_launch'],2))"
  done
done
