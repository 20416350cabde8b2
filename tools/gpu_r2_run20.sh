#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_kernels_gpu.py -k "implicit_gemm" -q --tb=short -p no:cacheprovider --timeout 120 --timeout-method=thread 2>&1 | tail -3
python tools/big_variant_bench.py 2>&1 | grep -v Warn | tail -3
for fl in 257 129; do
BBBP_FLASH_MIN_SEQ=$fl python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r2e_bench_fl$fl.json 2> gpurun_out/r2e_bench_fl$fl.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2e_bench_fl$fl.json'))
print('flash_min_seq=$fl', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], {k:(v['value'],v['e2e']) for k,v in d['by_precision'].items()})
PY
done
BBBP_FLASH_MIN_SEQ=129 timeout 900 python -m pytest tests/test_trained_parity_gpu.py -q --tb=short -p no:cacheprovider -s --timeout 600 --timeout-method=thread 2>&1 | grep -E "trained parity|passed|failed|FAILED|Error|assert" | tail -16
