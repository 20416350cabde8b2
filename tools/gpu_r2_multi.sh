#!/bin/bash
# multi-GPU round-2 measurements on ONE box with G GPUs (G = number visible): screening shards, DP training, H2D probe, bench
cd "$(dirname "$0")/.."
G=${G:-$(nvidia-smi -L | wc -l)}
N10=${N10:-10000000}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "${@:2}"; }
for n in 8 4 2; do
  [ $n -le $G ] || continue
  MODEL=morgan PRECISION=strict N=$N10 run $n tools/screen_10m.py 2>/dev/null | tail -1 | tee gpurun_out/r2_screen10m_morgan_strict_n$n.json
done
for n in 2 4 8; do
  [ $n -le $G ] || continue
  run $n tools/dp_train_bench.py 2>/dev/null | tail -1 | tee gpurun_out/r2_dp_train_n$n.json
  BATCH=256 STEPS=10 run $n tools/dp_train_bench.py 2>/dev/null | tail -1 | tee gpurun_out/r2_dp_train_b256_n$n.json
done
python tools/dp_train_bench.py 2>/dev/null | tail -1 | tee gpurun_out/r2_dp_train_n1.json
for n in 1 2 4 8; do
  [ $n -le $G ] || continue
  if [ $n -eq 1 ]; then python tools/h2d_probe.py 2>/dev/null | tail -1 | tee gpurun_out/r2_h2d_probe_n1.json
  else run $n tools/h2d_probe.py 2>/dev/null | tail -1 | tee gpurun_out/r2_h2d_probe_n$n.json; fi
done
if [ $G -ge 2 ]; then
  run $G bench.py --gpus $G --steps 8 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2_bench_n$G.json; head -c 600 gpurun_out/r2_bench_n$G.json; echo
fi
