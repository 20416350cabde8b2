#!/bin/bash
# Round-1 evidence refresh: default bench JSON, launch lists (bench + eager train steps), train-step device timeline.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.log 2>gpurun_out/bench_default.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench_default.log
CMD="python bench.py --steps 2 --warmup 3 --precision bf16 --groups 32 --no-cpu-baseline --no-train"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "bench launch list exit $?"
STEPS=4 python tools/train_steps.py > gpurun_out/train_plain.log 2>&1 && STEPS=4 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/train_launches.csv python tools/train_steps.py > gpurun_out/train_ncu.log 2>&1
echo "train launch list exit $?"
python tools/train_timeline.py 2>&1 | grep -v Warn | tail -30 > gpurun_out/train_timeline_summary.txt; cat gpurun_out/train_timeline_summary.txt | head -8
python tools/variant_bench.py 2>&1 | grep -v Warn | tail -12 | tee gpurun_out/variant_bench.log
timeout 600 python -m pytest -q --timeout 300 -p no:cacheprovider tests -m gpu -k "morgan_variant" 2>&1 | tail -3
