#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_kernels_gpu.py -k "background_referenced or pre_activation" -q --tb=short -p no:cacheprovider -s --timeout 120 --timeout-method=thread 2>&1 | grep -E "conv bg|passed|failed|FAILED|Error|assert" | tail -30
timeout 300 python tools/conv_bg_bench.py 2>&1 | tail -40
timeout 900 python -m pytest tests/test_trained_parity_gpu.py -q --tb=short -p no:cacheprovider -s --timeout 600 --timeout-method=thread 2>&1 | tee gpurun_out/r2_trained.log | grep -E "trained parity|passed|failed|FAILED|Error" | tail -30
