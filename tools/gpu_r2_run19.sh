#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_kernels_gpu.py -k "implicit_gemm or im2col or 64_output" -q --tb=short -p no:cacheprovider --timeout 120 --timeout-method=thread 2>&1 | tail -12
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_trained_parity_gpu.py -k "big" -q --tb=short -p no:cacheprovider -s --timeout 300 --timeout-method=thread 2>&1 | grep -E "trained parity|passed|failed|FAILED|Error|assert" | tail -12
python tools/big_variant_bench.py 2>&1 | grep -v Warn | tail -6
N=4096 python tools/big_variant_bench.py 2>&1 | grep -v Warn | tail -3
VARIANT=tcnn_big N=1024 PREC=bf16 python tools/infer_timeline.py 2>&1 | tail -12
