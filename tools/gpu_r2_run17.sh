#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_kernels_gpu.py -k "ffn_layernorm" -q --tb=short -p no:cacheprovider --timeout 120 --timeout-method=thread 2>&1 | tail -5
python tools/ffn_bench.py 2>&1 | tail -5
timeout 900 python -m pytest tests/test_model_gpu.py -k "c_host or c_program" -q --tb=short -p no:cacheprovider -s --timeout 300 --timeout-method=thread 2>&1 | grep -E "c host|passed|failed|FAILED|Error|assert" | tail -25
ROWS=16384 REPS=2 ncu --set full --clock-control none --import-source on -k regex:ffn_layernorm -s 3 -c 1 -o gpurun_out/r2_prof_ffn2 -f python tools/ffn_bench.py > gpurun_out/ncu_ffn.log 2>&1
