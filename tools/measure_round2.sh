#!/usr/bin/env bash
# GPU passes of round 2, in the order they were run (each: `gpurun -- bash tools/measure_round2.sh <pass>`;
# results land in gpurun_out/ and are summarised in profiles/r02_*).  `final` is what the driver runs at
# round end on one GPU.  Passes that tested kernel variants which were removed again are kept for the record
# of what was measured (profiles/README.md, "Round-2 experiments").
set -u
pass="${1:-final}"; shift || true
o=gpurun_out; mkdir -p $o
case "$pass" in
r2a)
  # Round-2 GPU pass A: GPU tests, driver-shaped bench lines (both arms), ncu captures of the FDL,
  # partition-MAC and direction-switching kernels.  Usage on the GPU box: bash tools/measure_round2.sh r2a
  ( time timeout 900 python -m pytest tests -m gpu -x -q ) > $o/r2a_pytest.log 2>&1
  tail -5 $o/r2a_pytest.log
  ( time timeout 400 python bench.py --steps 20 --warmup 5 ) > $o/r2a_bench.json 2> $o/r2a_bench.err
  tail -c 600 $o/r2a_bench.err
  timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > $o/r2a_bench_reference.json 2> $o/r2a_bench_reference.err
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:ols_fdl -s 24 -c 1 -f -o $o/prof_r2a_fdl \
      python tools/bench_fdl.py --streams 16384 --steps 4 --warmup 2 > $o/r2a_ncu_fdl.log 2>&1
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:sh_partition_mac -s 2 -c 1 -f -o $o/prof_r2a_mac \
      python tools/profile_run.py --workload c3 --streams 4096 --steps 4 > $o/r2a_ncu_mac.log 2>&1
  ls -la $o | grep r2a
  ;;
r2b)
  # Round-2 GPU pass B: the tensor-core partition MAC (parity first, then c3 timing and an ncu capture).
  ( time timeout 600 python -m pytest tests/test_gpu_partitioned.py tests/test_gpu_dropin.py -x -q -s ) > $o/r2b_pytest.log 2>&1
  tail -15 $o/r2b_pytest.log
  for tc in 1 0; do
    RTB_PART_TC=$tc timeout 300 python bench.py --workload c3 --streams 4096 --steps 50 --warmup 20 --no-cpu-baseline --no-extras \
       > $o/r2b_bench_c3_tc$tc.json 2> $o/r2b_bench_c3_tc$tc.err
    tail -c 300 $o/r2b_bench_c3_tc$tc.err
  done
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $o/r2b_launches_c3.csv \
      python tools/profile_run.py --workload c3 --streams 4096 --steps 5 > $o/r2b_ncu_list.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:sh_partition_tc -s 2 -c 1 -f -o $o/prof_r2b_pt \
      python tools/profile_run.py --workload c3 --streams 4096 --steps 4 > $o/r2b_ncu_pt.log 2>&1
  ls -la $o | grep r2b
  ;;
r2c)
  # Round-2 GPU pass C: whole GPU test-suite (incl. fused HPCF), c4 chain timing with the fused HPCF
  ( time timeout 1200 python -m pytest tests -m gpu -q -x ) > $o/r2c_pytest.log 2>&1
  tail -8 $o/r2c_pytest.log
  ls -la $o | grep r2c
  ;;
r2d)
  # Round-2 GPU pass D: restructured tensor-core partition MAC (parity, c3 timing, ncu), device SH transform test
  ( time timeout 600 python -m pytest tests/test_gpu_partitioned.py tests/test_gpu_api.py tests/test_gpu_hpcf_fused.py -x -q -s ) > $o/r2d_pytest.log 2>&1
  tail -6 $o/r2d_pytest.log
  timeout 300 python bench.py --workload c3 --streams 4096 --steps 50 --warmup 20 --no-cpu-baseline --no-extras \
       > $o/r2d_bench_c3.json 2> $o/r2d_bench_c3.err
  tail -c 300 $o/r2d_bench_c3.err
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $o/r2d_launches_c3.csv \
      python tools/profile_run.py --workload c3 --streams 4096 --steps 5 > $o/r2d_ncu_list.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"sh_partition_tc|sh_spectra_soa" -s 4 -c 2 -f -o $o/prof_r2d_pt \
      python tools/profile_run.py --workload c3 --streams 4096 --steps 4 > $o/r2d_ncu_pt.log 2>&1
  ls -la $o | grep r2d
  ;;
r2e)
  # Round-2 GPU pass E: folded pre-renderer (c4 chain), whole GPU test-suite, c4 chain timings
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
  ;;
r2f)
  # Round-2 GPU pass F (N GPUs): what limits the host-buffer (e2e) scaling.  Usage: bash tools/measure_round2.sh r2f N
  N=${1:-4}
  nvidia-smi topo -m > $o/r2f_topo_n$N.txt 2>&1
  lscpu | grep -i -E "numa|socket|model name|^cpu\(s\)" >> $o/r2f_topo_n$N.txt 2>&1
  ls /sys/devices/system/node/ >> $o/r2f_topo_n$N.txt 2>&1
  for b in "" "--bind"; do
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        tools/h2d_ranks.py $b > $o/r2f_h2d_n${N}${b}.json 2> $o/r2f_h2d_n${N}${b}.err
  done
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $N --steps 20 --warmup 5 > $o/r2f_bench_n$N.json 2> $o/r2f_bench_n$N.err
  tail -c 300 $o/r2f_bench_n$N.err
  ls -la $o | grep r2f
  ;;
r2g)
  # Round-2 GPU pass G: m-grouped render kernel for 32 / 64-sample blocks (tests, c4 timings)
  ( time timeout 900 python -m pytest tests -m gpu -q -x ) > $o/r2g_pytest.log 2>&1
  tail -5 $o/r2g_pytest.log
  timeout 300 python bench.py --workload c4 --streams 1024 --steps 200 --warmup 20 --no-cpu-baseline --no-extras > $o/r2g_bench_c4.json 2> $o/r2g_bench_c4.err
  RTB_RENDER_MG=0 timeout 300 python bench.py --workload c4 --streams 1024 --steps 200 --warmup 20 --no-cpu-baseline --no-extras > $o/r2g_bench_c4_termwise.json 2>> $o/r2g_bench_c4.err
  for S in 1024 4096; do
    timeout 300 python tools/bench_c4_chain.py --streams $S --steps 100 --fused > $o/r2g_c4_chain_fused_s$S.json 2> $o/r2g_c4_chain_fused_s$S.err
  done
  ls -la $o | grep r2g
  ;;
r2h)
  # Round-2 GPU pass H: whole GPU test-suite with the regression bars, MN-major TF32 probe (retry), fd_render ncu
  ( time timeout 900 python -m pytest tests -m gpu -q -x ) > $o/r2h_pytest.log 2>&1
  tail -5 $o/r2h_pytest.log
  timeout 60 ./build/tc_probe_mn > $o/r2h_tc_probe_mn.txt 2>&1
  cat $o/r2h_tc_probe_mn.txt
  timeout 120 python tools/profile_fd.py --steps 50 > $o/r2h_fd_render.json 2> $o/r2h_fd_render.err
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:fd_render -s 10 -c 1 -f -o $o/prof_r2h_fd \
      python tools/profile_fd.py --steps 3 > $o/r2h_ncu_fd.log 2>&1
  ls -la $o | grep r2h
  ;;
r2i)
  ( time timeout 600 python -m pytest tests/test_gpu_fd.py tests/test_gpu_large_blocks.py -q -x ) > $o/r2i_pytest.log 2>&1
  tail -4 $o/r2i_pytest.log
  timeout 120 python tools/profile_fd.py --steps 50 > $o/r2i_fd_render.json 2> $o/r2i_fd_render.err
  cat $o/r2i_fd_render.json
  timeout 120 python tools/profile_fd.py --steps 50 --taps 8192 > $o/r2i_fd_render_p32.json 2>> $o/r2i_fd_render.err
  cat $o/r2i_fd_render_p32.json
  ;;
r2j)
  ( time timeout 600 python -m pytest tests/test_gpu_partitioned.py -q -x ) > $o/r2j_pytest.log 2>&1
  tail -4 $o/r2j_pytest.log
  timeout 300 python bench.py --workload c3 --streams 4096 --steps 50 --warmup 20 --no-cpu-baseline --no-extras > $o/r2j_bench_c3.json 2> $o/r2j_bench_c3.err
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file $o/r2j_launches_c3.csv \
      python tools/profile_run.py --workload c3 --streams 4096 --steps 4 > $o/r2j_ncu_list.log 2>&1
  ;;
r2k)
  for d in 0 1 2 4 3 7; do
    RTB_PT_DEBUG=$d timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sh_partition_tc -c 3 --csv --log-file $o/r2k_dbg$d.csv \
      python tools/profile_run.py --workload c3 --streams 4096 --steps 4 > $o/r2k_dbg$d.log 2>&1
    echo "debug=$d $(grep sh_partition_tc $o/r2k_dbg$d.csv | awk -F'\",\"' '{print $NF}' | tr -d '\"' | tr '\n' ' ')"
  done
  ;;
r2l)
  ( time timeout 600 python -m pytest tests/test_gpu_partitioned.py -q -x -k "c3 or tc" ) > $o/r2l_pytest.log 2>&1
  tail -4 $o/r2l_pytest.log
  for d in 0 1 2 7; do
    RTB_PT_DEBUG=$d timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sh_partition_tc -c 3 --csv --log-file $o/r2l_dbg$d.csv \
      python tools/profile_run.py --workload c3 --streams 4096 --steps 4 > $o/r2l_dbg$d.log 2>&1
    echo "debug=$d $(grep sh_partition_tc $o/r2l_dbg$d.csv | awk -F'\",\"' '{print $NF}' | tr -d '\"' | tr '\n' ' ')"
  done
  timeout 300 python bench.py --workload c3 --streams 4096 --steps 50 --warmup 20 --no-cpu-baseline --no-extras > $o/r2l_bench_c3.json 2> $o/r2l_bench_c3.err
  python -c "import json; d=json.load(open('$o/r2l_bench_c3.json')); print('c3', d['value'], d['ms_per_step'])"
  ;;
r2m)
  # Round-2 GPU pass M: dual-signal render kernel for small batches (parity, then timing per batch size)
  ( time timeout 900 python -m pytest tests -m gpu -q -x ) > $o/r2m_pytest.log 2>&1
  tail -4 $o/r2m_pytest.log
  for S in 512 1024 1184 1536 2048; do
    for d in 0 1; do
      RTB_RENDER_DUAL=$d timeout 200 python bench.py --streams $S --steps 200 --warmup 20 --no-cpu-baseline --no-extras > $o/r2m_bench_s${S}_dual$d.json 2> $o/r2m_err.txt
      python -c "import json; d=json.load(open('$o/r2m_bench_s${S}_dual$d.json')); print('S=$S dual=$d ms', round(d['ms_per_step'],5), 'p99', round(d['p99_block_latency_ms'],5), 'render us', round(d['roofline']['us_per_launch'],2))"
    done
  done
  ;;
r2n)
  ( time timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_gpu_sh_chain.py tests/test_gpu_hpcf_fused.py -q -x ) > $o/r2n_pytest.log 2>&1
  tail -3 $o/r2n_pytest.log
  for S in 512 1024 2048; do
    for mg in default 1; do
      if [ $mg = default ]; then unset RTB_RENDER_MG; else export RTB_RENDER_MG=1; fi
      timeout 200 python bench.py --workload c4 --streams $S --steps 200 --warmup 20 --no-cpu-baseline --no-extras > $o/r2n_c4_s${S}_mg$mg.json 2> $o/r2n_err.txt
      python -c "import json; d=json.load(open('$o/r2n_c4_s${S}_mg$mg.json')); print('c4 S=$S mg=$mg ms', round(d['ms_per_step'],5), d['roofline']['kernel'], round(d['roofline']['us_per_launch'],2))"
    done
  done
  ;;
final)
  # What the driver runs at round end, on one GPU: GPU tests, smoke(), both bench arms.
  ( time timeout 1200 python -m pytest tests -m gpu -q -x ) > $o/final_pytest.log 2>&1
  tail -4 $o/final_pytest.log
  ( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $o/final_smoke.log 2>&1
  tail -3 $o/final_smoke.log
  ( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $o/final_bench_reference.json 2> $o/final_bench_reference.err
  ( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $o/final_bench.json 2> $o/final_bench.err
  tail -3 $o/final_bench.err
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/final_launches_bench_c2.csv \
      python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $o/final_ncu_list.log 2>&1
  timeout 200 python tools/profile_fd.py --streams 4096 --steps 200 > $o/final_fd_split_s4096.json 2>> $o/final_bench.err
  timeout 200 python tools/profile_fd.py --streams 4096 --taps 8192 --steps 100 > $o/final_fd_split_s4096_p32.json 2>> $o/final_bench.err
  RTB_FD_SPLIT=0 timeout 200 python tools/profile_fd.py --streams 4096 --steps 200 > $o/final_fd_fused_s4096.json 2>> $o/final_bench.err
  tail -qn1 $o/final_fd_*.json | cut -c1-260
  for S in 1024 4096; do timeout 300 python tools/bench_c4_chain.py --fused --streams $S > $o/final_c4_chain_s${S}_fused.json 2>> $o/final_bench.err; tail -n1 $o/final_c4_chain_s${S}_fused.json | cut -c1-200; done
  ls -la $o | grep final
  ;;
*) echo "unknown pass $pass"; exit 2 ;;
esac
