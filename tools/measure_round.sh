#!/usr/bin/env bash
# One-GPU measurement pass behind profiles/: bench lines, sweeps, launch list, full ncu captures.
# Usage (on the GPU box): bash tools/measure_round.sh <tag>     -> gpurun_out/<tag>_*
set -u
tag="${1:-r01b}"
o=gpurun_out
mkdir -p $o
run() { timeout "$1" "${@:2}"; }
run 400 python bench.py > $o/${tag}_bench_c2_s1024.json 2> $o/${tag}_bench_c2_s1024.err
for S in 4096 10240 16384; do
  run 200 python bench.py --streams $S --steps 500 --warmup 50 --no-cpu-baseline > $o/${tag}_bench_c2_s$S.json 2> $o/${tag}_bench_c2_s$S.err
done
run 200 python bench.py --workload c5 --streams 8192 --steps 200 --warmup 20 --no-cpu-baseline > $o/${tag}_bench_c5_s8192.json 2> $o/${tag}_bench_c5.err
run 200 python bench.py --workload c3 --streams 4096 --steps 100 --warmup 10 --no-cpu-baseline > $o/${tag}_bench_c3_s4096.json 2> $o/${tag}_bench_c3.err
run 200 python bench.py --workload c4 --streams 1024 --steps 200 --warmup 20 --no-cpu-baseline > $o/${tag}_bench_c4_s1024.json 2> $o/${tag}_bench_c4.err
run 400 python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference.json 2> $o/${tag}_bench_reference.err
run 100 python tools/bench_fdl.py --streams 16384 > $o/${tag}_fdl_s16384.json 2> $o/${tag}_fdl.err
run 100 python tools/bench_fdl.py --streams 16384 --taps 1024 --block 64 > $o/${tag}_fdl_hpcf_s16384.json 2>> $o/${tag}_fdl.err
# launch list of the default bench command (per-launch times are cold-cache and serialised)
run 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/${tag}_launches_bench_c2.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $o/${tag}_ncu_list.log 2>&1
# full captures of both kernels of the chain
run 300 ncu --set full --clock-control none --import-source on -k regex:sh_ -s 2 -c 2 -f -o $o/prof_${tag}_c2_s1024 \
    python tools/profile_run.py --streams 1024 --steps 4 > $o/${tag}_ncu_full_1024.log 2>&1
run 300 ncu --set full --clock-control none --import-source on -k regex:sh_ -s 2 -c 2 -f -o $o/prof_${tag}_c2_s10240 \
    python tools/profile_run.py --streams 10240 --steps 4 > $o/${tag}_ncu_full_10240.log 2>&1
ls -la $o | grep ${tag} | head -40
