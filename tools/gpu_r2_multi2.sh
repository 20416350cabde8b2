#!/bin/bash
# G-GPU box: bench.py --gpus G (strict, sparse e2e) and the C ABI's NCCL exchanges on G ranks
cd "$(dirname "$0")/.."
G=${G:-$(nvidia-smi -L | wc -l)}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "${@:2}"; }
run $G bench.py --gpus $G --steps 8 --warmup 3 2>gpurun_out/r2m_bench$G.err | tail -1 > gpurun_out/r02_bench_n$G.json
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_n$G.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'], 'dense', d['e2e_dense_u8']['value'], {k:(v['value'],v['e2e']) for k,v in d['by_precision'].items()})
PY
run $G tools/comm_check.py 2>/dev/null | tail -1 | tee gpurun_out/r02_comm_check_n$G.json
