#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 --timeout-method=thread -k "unpack or sparse or conv or packed or screening or smoke or predict_from_host" 2>&1 | grep -E "passed|failed|FAILED|Error|assert" | tail -20
python tools/hbm_kernels.py > gpurun_out/r2_hbm.log 2>&1; cp gpurun_out/hbm_kernels.txt gpurun_out/r02_hbm_kernels.txt; grep -E "unpack|packed|sparse" gpurun_out/hbm_kernels.txt
timeout 300 python tools/conv_bg_bench.py 2>&1 | grep -E "BG|one pass|pairs" | tail -16
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r2_bench_bg.json 2> gpurun_out/r2_bench_bg.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_bg.json'))
print({k:d[k] for k in ('value','ms_per_step','dtype','gpu_launches')}, 'e2e', d['e2e']['value'], 'conv2 ms', d['roofline']['launch_ms'], 'conv1 ms', d['roofline']['conv1_launch_ms'], d['by_precision'])
PY
tail -3 gpurun_out/r2_bench_bg.err
