#!/bin/bash
# Training-path GPU stage: graphed-step tests, eager vs graph timing, device timeline of one replay.
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 300 --timeout-method thread -p no:cacheprovider"
timeout 900 $PT tests -m gpu -k "graphed or attention or dropout or adamw or train or gemm_f32 or layernorm or conv_relu_pool" > gpurun_out/pytest_train.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_train.log
timeout 600 python tools/train_bench.py > gpurun_out/train_bench.log 2>&1; echo "train_bench exit $?"; tail -1 gpurun_out/train_bench.log
timeout 300 python tools/train_timeline.py 2>&1 | grep -v Warn | tail -32
B=256 timeout 300 python tools/train_timeline.py 2>&1 | grep -v Warn | tail -22
