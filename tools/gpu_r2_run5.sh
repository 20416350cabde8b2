#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s --timeout 600 --timeout-method=thread 2>&1 | tee gpurun_out/r2_tests.log | grep -E "tc training|passed|failed|FAILED|Error" | tail -40
PRECS=fp32,bf16,fp16 BATCHES=32,64,256 python tools/train_bench.py 2>&1 | tail -12 | tee gpurun_out/r2_train_bench.txt
MODEL=morgan PRECISION=strict python tools/screen_10m.py 2>&1 | tail -1 | tee gpurun_out/r2_screen10m_morgan_strict_n1.json
python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','dtype','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e_streaming']['value'], d['by_precision'], d['roofline']['launch_ms'], d['roofline']['conv1_launch_ms'], d['train_step'])
PY
python tools/ncu_targets.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"attention_flash|gemm_bf16|conv3x3_umma|add_layernorm" -s 20 -c 24 -o gpurun_out/r2_prof python tools/ncu_targets.py > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log; ls -la gpurun_out/r2_prof.ncu-rep
