#!/bin/bash
# re-entry check: full GPU suite, smoke, default bench (both arms), launch list of the bench
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 --timeout-method=thread 2>&1 | tee gpurun_out/r2b_tests.log | grep -E "passed|failed|FAILED|Error" | tail -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','dtype','gpu_launches')}, 'e2e', d['e2e']['value'], 'conv2 ms', d['roofline']['launch_ms'], 'conv1 ms', d['roofline']['conv1_launch_ms'], d['by_precision'], d.get('train_step'))
PY
tail -3 gpurun_out/r2b_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_bench_ref.json 2>/dev/null; head -c 600 gpurun_out/r2b_bench_ref.json
