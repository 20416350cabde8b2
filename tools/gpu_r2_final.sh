#!/bin/bash
# round-2 evidence run: both bench arms, launch list of the bench command, full ncu captures of the dominant kernels (each ncu
# pass only after the same command exited 0 without ncu).  Reports are summarised ON the box (gpurun_out/ is capped at 64 MiB).
cd "$(dirname "$0")/.."
if [ "$1" == "tests" ]; then
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 --timeout-method=thread 2>&1 | tee gpurun_out/r2f_tests.log | grep -E "passed|failed|FAILED|Error" | tail -20
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
fi
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_n1.json 2>/dev/null
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r2f_bench.err && python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','dtype','gpu_launches')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['kernel'][:5], d['roofline']['frac'], d['roofline']['launch_ms'], d['roofline_second']['kernel'][:5], d['roofline_second']['frac'], d['roofline_second']['launch_ms'], {k:(v['value'],v['e2e']) for k,v in d['by_precision'].items()}, d.get('clocks'))
PY
python tools/hbm_kernels.py > gpurun_out/r2f_hbm.log 2>&1; cp gpurun_out/hbm_kernels.txt gpurun_out/r02_hbm_kernels.txt 2>/dev/null
python tools/ffn_bench.py 2>&1 | tail -4 > gpurun_out/r02_ffn_bench.txt
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --also "" > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --also "" > gpurun_out/ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_bench.csv > gpurun_out/r02_launches_bench_summary.txt 2>&1; head -12 gpurun_out/r02_launches_bench_summary.txt
python tools/ncu_targets.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"attention_flash|ffn_layernorm|conv3x3_umma" -s 14 -c 14 -o /tmp/r02_prof_a -f python tools/ncu_targets.py > gpurun_out/ncu_a.log 2>&1
python tools/ncu_summary.py /tmp/r02_prof_a.ncu-rep > gpurun_out/r02_ncu_conv_ffn_flash.txt 2>&1
ncu --set full --clock-control none -k regex:"conv3x3_umma" -s 4 -c 2 -o /tmp/r02_prof_c -f python tools/ncu_targets.py > gpurun_out/ncu_c.log 2>&1
python tools/ncu_summary.py /tmp/r02_prof_c.ncu-rep > gpurun_out/r02_ncu_conv_fp32_planes.txt 2>&1
ncu --set full --clock-control none -k regex:"gemm_bf16_kernel" -s 60 -c 12 -o /tmp/r02_prof_b -f python tools/ncu_targets.py > gpurun_out/ncu_b.log 2>&1
python tools/ncu_summary.py /tmp/r02_prof_b.ncu-rep > gpurun_out/r02_ncu_gemm.txt 2>&1
ls -la /tmp/*.ncu-rep; cut -c1-220 gpurun_out/r02_ncu_conv_ffn_flash.txt | head -20
