#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 --timeout-method=thread 2>&1 | tee gpurun_out/r2_tests.log | grep -E "passed|failed|FAILED|Error" | tail -30
timeout 600 python tests/strict_error_budget.py 2>&1 | grep -v Warn | tee gpurun_out/r2_strict_error_budget.txt | tail -40
python tools/tmem_probe.py 2>&1 | tee gpurun_out/r2_tmem_probe.txt
python tools/hbm_kernels.py > gpurun_out/r2_hbm.log 2>&1; cat gpurun_out/hbm_kernels.txt
PRECISION=strict N=16384 python tools/e2e_sweep.py 2>&1 | tee gpurun_out/r2_e2e_sweep_strict.txt | tail -8
PRECISION=bf16 python tools/batch_sweep.py 2>&1 | tee gpurun_out/r2_batch_sweep_bf16.txt | tail -18; cp gpurun_out/batch_sweep.json gpurun_out/r2_batch_sweep_bf16.json
PRECISION=strict python tools/batch_sweep.py 2>&1 | tee gpurun_out/r2_batch_sweep_strict.txt | tail -18; cp gpurun_out/batch_sweep.json gpurun_out/r2_batch_sweep_strict.json
MODEL=morgan PRECISION=strict python tools/screen_10m.py 2>&1 | tail -1 | tee gpurun_out/r2_screen10m_morgan_strict_n1.json
