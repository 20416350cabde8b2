#!/bin/bash
# check of the committed state: full GPU suite, smoke, (with "bench") both bench arms
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 --timeout-method=thread 2>&1 | tee gpurun_out/r2v_tests.log | grep -E "passed|failed|FAILED|Error" | tail -20
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
[ "$1" == "bench" ] || exit 0
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_n1.json 2>/dev/null
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r2v_bench.err && python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','dtype','gpu_launches')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['kernel'][:5], d['roofline']['frac'], d['roofline']['launch_ms'], d['roofline_second']['kernel'][:5], d['roofline_second']['frac'], d['roofline_second']['launch_ms'], {k:(v['value'],v['e2e']) for k,v in d['by_precision'].items()}, d.get('clocks'), d['train_step'])
PY
