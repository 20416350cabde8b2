#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_kernels_gpu.py -k "background_referenced or pre_activation or conv_stack or conv1 or conv3x3" -q --tb=short -p no:cacheprovider -s --timeout 120 --timeout-method=thread 2>&1 | grep -E "passed|failed|FAILED|Error|assert" | tail -30
timeout 300 python tools/conv_bg_bench.py 2>&1 | tail -24
timeout 900 python -m pytest tests/test_trained_parity_gpu.py -q --tb=short -p no:cacheprovider -s --timeout 600 --timeout-method=thread 2>&1 | tee gpurun_out/r2_trained.log | grep -E "trained parity.*strict|passed|failed|FAILED|Error" | tail -30
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r2_bench_bg.json 2> gpurun_out/r2_bench_bg.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_bg.json'))
print({k:d[k] for k in ('value','ms_per_step','dtype','gpu_launches')}, 'e2e', d['e2e']['value'], 'conv2 ms', d['roofline']['launch_ms'], 'conv1 ms', d['roofline']['conv1_launch_ms'], d['by_precision'])
PY
tail -3 gpurun_out/r2_bench_bg.err
