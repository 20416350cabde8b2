#!/bin/bash
# 4 ranks share one host bridge on this pool's boxes: the e2e chunk schedule with and without the long second chunk
cd "$(dirname "$0")/.."
G=${G:-4}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "${@:2}"; }
for sched in 2048,6144,8192 2048,14336 4096,12288; do
BBBP_E2E_SCHEDULE=$sched run $G bench.py --gpus $G --steps 8 --warmup 3 --also "" 2>/dev/null | tail -1 > gpurun_out/r02_bench_n${G}_$sched.json
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_n${G}_$sched.json'))
print('$sched', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
PY
done
