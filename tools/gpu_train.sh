#!/bin/bash
# Training-path GPU stage: graphed-step tests, eager vs graph timing, device timeline of one replay, ncu of the chain kernels.
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 300 --timeout-method thread -p no:cacheprovider"
timeout 900 $PT tests -m gpu -k "graphed or attention or dropout or adamw or train or gemm_f32 or layernorm or conv_relu_pool" > gpurun_out/pytest_train.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_train.log
timeout 600 python tools/train_bench.py > gpurun_out/train_bench.log 2>&1; echo "train_bench exit $?"; tail -1 gpurun_out/train_bench.log
timeout 300 python tools/train_timeline.py 2>&1 | grep -v Warn | tail -32
STEPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_f32_skinny|attention_short" -s 40 -c 12 -o gpurun_out/prof_train_chain -f python tools/train_steps.py > gpurun_out/ncu_train_chain.log 2>&1; echo "ncu exit $?"
