#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --precision bf16 --groups 32 --no-cpu-baseline --no-train"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv3x3_umma_kernel -s 4 -c 2 -o gpurun_out/prof_conv -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
