#!/bin/bash
# One gpurun call: staged so that a failure in one stage (e.g. a faulting kernel poisoning its CUDA context)
# does not hide the others.  Every stage has its own timeout and log under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n "${TAIL:-12}" gpurun_out/$name.log; return $rc; }
PT="python -m pytest -q --timeout 120 --timeout-method thread -p no:cacheprovider"
TMO=1500 run pytest_fp32   $PT tests -m gpu -k "not bf16 and not tcgen05"
TMO=600  run pytest_tc     $PT tests -m gpu -k "bf16 or tcgen05" || {
  BBBP_CONV_SWAP_DESC=1 TMO=300 run pytest_conv_swapped $PT tests/test_kernels_gpu.py -m gpu -k "conv1_tcgen05 or conv2_tcgen05 or localises"; }
TMO=300  run smoke         python -c "import __graft_entry__ as g; g.smoke()"
TMO=600  run bench_fp32    python bench.py --steps 3 --warmup 3 --precision fp32 --groups 8
TMO=600  run bench_bf16    python bench.py --steps 5 --warmup 3 --precision bf16 --groups 32 --no-cpu-baseline
