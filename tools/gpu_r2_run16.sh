#!/bin/bash
cd "$(dirname "$0")/.."
python tools/ffn_bench.py 2>&1 | tail -5
ROWS=16384 REPS=2 ncu --set full --clock-control none --import-source on -k regex:ffn_layernorm -s 3 -c 1 -o gpurun_out/r2_prof_ffn -f python tools/ffn_bench.py > gpurun_out/ncu_ffn.log 2>&1
tail -3 gpurun_out/ncu_ffn.log; ls -la gpurun_out/r2_prof_ffn.ncu-rep
