#!/bin/bash
# round-2 evidence run: full GPU suite, smoke, both bench arms, launch list of the bench command, full ncu captures of the
# dominant kernels (each ncu pass only after the same command exited 0 without ncu)
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 --timeout-method=thread 2>&1 | tee gpurun_out/r2f_tests.log | grep -E "passed|failed|FAILED|Error" | tail -20
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_n1.json 2>/dev/null
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r2f_bench.err && python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','dtype','gpu_launches')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['launch_ms'], d['roofline']['conv1_launch_ms'], {k:(v['value'],v['e2e']) for k,v in d['by_precision'].items()}, d.get('train_step'), d.get('clocks'))
PY
python tools/hbm_kernels.py > gpurun_out/r2f_hbm.log 2>&1; cp gpurun_out/hbm_kernels.txt gpurun_out/r02_hbm_kernels.txt 2>/dev/null
python tools/ffn_bench.py 2>&1 | tail -4 > gpurun_out/r02_ffn_bench.txt; cat gpurun_out/r02_ffn_bench.txt
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_launches.log 2>&1
python tools/ncu_targets.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"attention_flash|ffn_layernorm|conv3x3_umma|gemm_bf16_kernel<128" -s 30 -c 40 -o gpurun_out/r02_prof_final -f python tools/ncu_targets.py > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log; ls -la gpurun_out/*.ncu-rep
