#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 --timeout-method=thread 2>&1 | tee gpurun_out/r2d_tests.log | grep -E "passed|failed|FAILED|Error" | tail -20
SPARSE=1 SCHEDULES=1 PRECISION=strict N=16384 python tools/e2e_sweep.py 2>&1 | grep -v Warn | tee gpurun_out/r02_e2e_sweep_sparse_strict.txt | tail -14
SPARSE=1 SCHEDULES=only PRECISION=fp16 N=16384 python tools/e2e_sweep.py 2>&1 | grep -v Warn | tee gpurun_out/r02_e2e_sweep_sparse_fp16.txt | tail -8
timeout 300 python tools/conv_bg_bench.py 2>&1 | grep -E "BG|one pass|pairs|exact" | tail -16
