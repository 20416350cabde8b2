#!/bin/bash
# round-2 GPU check: full GPU test suite (all failures listed), smoke, bench, launch list of the strict bench
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s 2>&1 | tee gpurun_out/r2_tests.log | grep -E "trained parity|conv strict|passed|failed|FAILED|Error" | tail -60
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 3000 gpurun_out/r2_bench.json
python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline --also "" > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_strict.csv \
    python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline --also "" > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
