#!/usr/bin/env bash
# A/B timing of the tcgen05 analysis kernel variants (run on the GPU box); every command under its own timeout
T="timeout 90"
timeout 200 python -m pytest tests/test_gpu_sh_chain.py tests/test_gpu_variants.py tests/test_gpu_large_blocks.py -x -q -m gpu 2>&1 | tail -3
$T python tools/kernel_times.py --workload c5 --streams 8192
RTB_TC_HALVES=1 $T python tools/kernel_times.py --workload c5 --streams 8192
$T python tools/kernel_times.py --workload c4 --streams 1024
RTB_ANALYSIS_TC=1 $T python tools/kernel_times.py --workload c2 --streams 10240
RTB_ANALYSIS_TC=1 RTB_TC_HALVES=1 $T python tools/kernel_times.py --workload c2 --streams 10240
RTB_ANALYSIS_TC=1 $T python tools/kernel_times.py --workload c2 --streams 1024
RTB_ANALYSIS_TC=1 RTB_TC_HALVES=1 $T python tools/kernel_times.py --workload c2 --streams 1024
export RTB_LIB_PATH=$PWD/build/lib_tctrace/libretisar_b200.so
$T python tools/tc_trace.py c5 8192
RTB_ANALYSIS_TC=1 $T python tools/tc_trace.py c2 10240
