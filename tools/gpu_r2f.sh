#!/usr/bin/env bash
# Round-2 GPU pass F (N GPUs): what limits the host-buffer (e2e) scaling.  Usage: bash tools/gpu_r2f.sh N
set -u
N=${1:-4}
o=gpurun_out
mkdir -p $o
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
