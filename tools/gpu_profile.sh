#!/bin/bash
# ncu evidence for the bench command: (1) launch list with per-launch device time, (2) full-set capture of the
# tcgen05 conv kernels.  Each ncu run is preceded by the same command exiting 0 without ncu.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --precision bf16 --groups 32 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv3x3_umma_kernel -s 4 -c 2 -o gpurun_out/prof_conv -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/plain.log
