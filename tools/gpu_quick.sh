#!/bin/bash
# quick iteration: tensor-core tests + bf16 bench
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 120 --timeout-method thread -p no:cacheprovider"
timeout 600 $PT tests -m gpu -k "bf16 or tcgen05 or ${EXTRA_K:-tcgen05}" > gpurun_out/quick_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/quick_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --precision bf16 --groups 32 --no-cpu-baseline > gpurun_out/quick_bench.log 2>&1; echo "bench exit $?"; tail -2 gpurun_out/quick_bench.log
