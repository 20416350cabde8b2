#!/bin/bash
# fused FFN + LayerNorm kernel, exact-integer first layer, C host: new tests first (bounded), then A/B bench
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_kernels_gpu.py -k "ffn_layernorm or background_referenced" -q --tb=short -p no:cacheprovider --timeout 120 --timeout-method=thread 2>&1 | tail -25
timeout 900 python -m pytest tests/test_model_gpu.py -k "c_host or c_program or comm_entry or wide_attention" -q --tb=short -p no:cacheprovider -s --timeout 300 --timeout-method=thread 2>&1 | grep -E "c host|passed|failed|FAILED|Error|assert" | tail -25
timeout 900 python -m pytest tests/test_trained_parity_gpu.py -q --tb=short -p no:cacheprovider -s --timeout 600 --timeout-method=thread 2>&1 | grep -E "trained parity|passed|failed|FAILED|Error|assert" | tail -25
for ffn in 1 0; do
BBBP_FUSED_FFN=$ffn python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r2c_bench_ffn$ffn.json 2> gpurun_out/r2c_bench_ffn$ffn.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2c_bench_ffn$ffn.json'))
print('fused_ffn=$ffn', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'conv2 ms', d['roofline']['launch_ms'], 'conv1 ms', d['roofline']['conv1_launch_ms'], {k:(v['value'],v['e2e']) for k,v in d['by_precision'].items()})
PY
done
PREC=strict N=16384 python tools/infer_timeline.py 2>&1 | tail -16
