#!/bin/bash
# background-referenced strict mode: kernel tests, trained parity in the three strict forms, error budget, bench
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_kernels_gpu.py -k "background_referenced or pre_activation or conv_stack" -q --tb=short -p no:cacheprovider -s --timeout 120 --timeout-method=thread 2>&1 | grep -E "conv bg|conv strict|passed|failed|FAILED|Error|assert" | tail -40
for cfg in "1 1" "1 0" "0 1"; do
  set -- $cfg
  echo "== BBBP_STRICT_BACKGROUND=$1 BBBP_STRICT_CONV1_SPLIT=$2"
  BBBP_STRICT_BACKGROUND=$1 BBBP_STRICT_CONV1_SPLIT=$2 timeout 900 python -m pytest tests/test_trained_parity_gpu.py -q --tb=short -p no:cacheprovider -s --timeout 600 --timeout-method=thread 2>&1 | grep -E "trained parity.*strict|passed|failed|FAILED|Error" | tail -12
done
for cfg in "1 1" "1 0" "0 1"; do
  set -- $cfg
  echo "== bench BBBP_STRICT_BACKGROUND=$1 BBBP_STRICT_CONV1_SPLIT=$2"
  BBBP_STRICT_BACKGROUND=$1 BBBP_STRICT_CONV1_SPLIT=$2 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-train --also "" > gpurun_out/r2_bench_bg_$1$2.json 2> gpurun_out/r2_bench_bg.err
  python - "$1$2" <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/r2_bench_bg_{sys.argv[1]}.json'))
print({k:d[k] for k in ('value','ms_per_step','dtype','gpu_launches')}, 'e2e', d['e2e']['value'], 'conv2 ms', d['roofline']['launch_ms'], 'conv1 ms', d['roofline']['conv1_launch_ms'])
PY
done
tail -5 gpurun_out/r2_bench_bg.err
