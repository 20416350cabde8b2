#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s --timeout 600 --timeout-method=thread -k "unpack or sparse or tensor_core_training or screening_on_packed or packed_bits" 2>&1 | grep -E "tc training|passed|failed|FAILED|Error" | tail -20
python tools/hbm_kernels.py > gpurun_out/r2_hbm.log 2>&1; cp gpurun_out/hbm_kernels.txt gpurun_out/r02_hbm_kernels.txt; cat gpurun_out/hbm_kernels.txt
SPARSE=1 PRECISION=strict N=16384 python tools/e2e_sweep.py 2>&1 | tee gpurun_out/r02_e2e_sweep_sparse_strict.txt | tail -8
SPARSE=1 PRECISION=fp16 N=16384 python tools/e2e_sweep.py 2>&1 | tee gpurun_out/r02_e2e_sweep_sparse_fp16.txt | tail -8
