#!/bin/bash
# round-2 GPU check: new kernels first (bounded), full GPU test suite (all failures listed), smoke, bench
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_kernels_gpu.py -k "flash" -q --tb=short -p no:cacheprovider --timeout 120 --timeout-method=thread 2>&1 | tail -15
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s --timeout 600 --timeout-method=thread 2>&1 | tee gpurun_out/r2_tests.log | grep -E "trained parity|passed|failed|FAILED|Error" | tail -40
timeout 600 python tests/strict_error_budget.py 2>&1 | grep -v Warn | tee gpurun_out/r2_strict_error_budget.txt | tail -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; head -c 1200 gpurun_out/r2_bench.json; echo; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','dtype','gpu_launches')}, d['e2e']['value'], d['by_precision'], d['roofline']['launch_ms'], d['roofline']['conv1_launch_ms'])
PY
