#!/usr/bin/env bash
set -u
o=gpurun_out
mkdir -p $o
( time timeout 600 python -m pytest tests/test_gpu_fd.py tests/test_gpu_large_blocks.py -q -x ) > $o/r2i_pytest.log 2>&1
tail -4 $o/r2i_pytest.log
timeout 120 python tools/profile_fd.py --steps 50 > $o/r2i_fd_render.json 2> $o/r2i_fd_render.err
cat $o/r2i_fd_render.json
timeout 120 python tools/profile_fd.py --steps 50 --taps 8192 > $o/r2i_fd_render_p32.json 2>> $o/r2i_fd_render.err
cat $o/r2i_fd_render_p32.json
