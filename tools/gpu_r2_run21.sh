#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_kernels_gpu.py -k "ffn_layernorm or flash" -q --tb=short -p no:cacheprovider --timeout 120 --timeout-method=thread 2>&1 | tail -15
python tools/ffn_bench.py 2>&1 | tail -4
timeout 900 python -m pytest tests/test_model_gpu.py -k "c_host or c_program or wide_attention or eval_forward" -q --tb=short -p no:cacheprovider --timeout 300 --timeout-method=thread 2>&1 | tail -8
for tail in 1 0; do
BBBP_FUSED_ATTENTION_TAIL=$tail python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r2g_bench_t$tail.json 2> gpurun_out/r2g_bench_t$tail.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2g_bench_t$tail.json'))
print('fused_tail=$tail', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], {k:(v['value'],v['e2e']) for k,v in d['by_precision'].items()})
PY
done
