#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s --timeout 600 --timeout-method=thread 2>&1 | tee gpurun_out/r2_tests.log | grep -E "tc training|passed|failed|FAILED|Error" | tail -40
python tools/pca_bench.py 2>&1 | tail -6 | tee gpurun_out/r02_pca_bench.txt
python tools/hbm_kernels.py > gpurun_out/r2_hbm.log 2>&1; cp gpurun_out/hbm_kernels.txt gpurun_out/r02_hbm_kernels.txt; cat gpurun_out/hbm_kernels.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 8 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r2_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','dtype','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], 'dense', d['e2e_dense_u8']['value'], 'stream', d['e2e_streaming']['value'], d['by_precision'], d['roofline']['launch_ms'], d['roofline']['conv1_launch_ms'], d['train_step'])
PY
python tools/ncu_targets.py > gpurun_out/plain.log 2>&1 && {
ncu --set full --clock-control none --import-source on -k regex:"attention_flash" -c 2 -o gpurun_out/r02_prof_flash python tools/ncu_targets.py > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv3x3_umma" -c 4 -o gpurun_out/r02_prof_conv python tools/ncu_targets.py > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none -k regex:"gemm_bf16" -s 8 -c 10 -o gpurun_out/r02_prof_gemm python tools/ncu_targets.py > gpurun_out/ncu3.log 2>&1
}
tail -2 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log; ls -la gpurun_out/*.ncu-rep
