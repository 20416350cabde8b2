#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s --timeout 600 --timeout-method=thread 2>&1 | tee gpurun_out/r2_tests.log | grep -E "trained parity|passed|failed|FAILED|Error" | tail -40
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/conv_bg_bench.py 2>&1 | grep -E "BG|one pass" | tail -12
PREC=strict N=16384 python tools/infer_timeline.py 2>&1 | tail -24
cp gpurun_out/infer_timeline_tcnn_167.txt gpurun_out/r02_infer_step_timeline_strict_16384.txt
python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','dtype','gpu_launches')}, 'e2e', d['e2e']['value'], 'conv2 ms', d['roofline']['launch_ms'], 'conv1 ms', d['roofline']['conv1_launch_ms'], d['by_precision'], d.get('train_step'))
PY
tail -3 gpurun_out/r2_bench.err
