"""Data-parallel training on NCCL hardware (BASELINE configs[1] "reported, not forced"): the reference loop body
(20250113.py:186-191) on R replicas with the flat gradient buffer averaged bucket by bucket during backward
(bbbp_b200.FlatGradients), fused AdamW replicated.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 --master-port P tools/dp_train_bench.py

Times, with CUDA events and the max over ranks: the step without communication (world 1 semantics), the step with the
overlapped bucketed all-reduce, the same step with ONE all-reduce after backward (no overlap), and the bare all-reduce of
the 54 MB buffer.  Rank 0 prints one JSON line; replicas must stay bit-identical (parameter checksum compared across ranks)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bbbp_b200

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
BATCH = int(os.environ.get("BATCH", 32)); STEPS = int(os.environ.get("STEPS", 30))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = bbbp_b200.MixedInputModel(167, 128).to(dev)
bbbp_b200.zero_dropout(model)
model.train()
opt = bbbp_b200.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)
crit = bbbp_b200.MSELoss()
g = torch.Generator(device=dev).manual_seed(100 + rank)                 # every rank its own micro-batch
fp = torch.randn(BATCH, 167, generator=g, device=dev); img = torch.randn(BATCH, 3 * 128 * 128, generator=g, device=dev)
y = torch.randn(BATCH, generator=g, device=dev) * 0.75 - 0.1


def timed(fn, steps=STEPS):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


flat = bbbp_b200.FlatGradients(model.parameters(), bucket_bytes=int(os.environ.get("BUCKET_MB", 16)) << 20)


def step_overlapped():
    flat.zero()
    crit(model(fp, img).squeeze(), y).backward()
    flat.synchronize()
    opt.step()


hooks, flat._hooks = flat._hooks, []                                     # the same buffer without the overlap machinery
for h in hooks:
    h.remove()


def step_local():
    flat.zero()
    crit(model(fp, img).squeeze(), y).backward()
    opt.step()


def step_one_allreduce():
    flat.zero()
    crit(model(fp, img).squeeze(), y).backward()
    if world > 1:
        dist.all_reduce(flat.flat, op=dist.ReduceOp.AVG)
    opt.step()


res = {"workload": f"DP training, MixedInputModel(167,128) fp32 kernels, batch {BATCH} per rank, AdamW", "n_gpus": world,
       "grad_buffer_MB": flat.flat.numel() * 4 / 1e6, "buckets": len(flat.buckets)}
res["step_ms_no_comm"] = timed(step_local)
res["step_ms_one_allreduce_after_backward"] = timed(step_one_allreduce)
if world > 1:
    res["allreduce_ms_alone"] = timed(lambda: dist.all_reduce(flat.flat, op=dist.ReduceOp.AVG), 20)
    res["allreduce_bus_GBs"] = 2 * (world - 1) / world * flat.flat.numel() * 4 / (res["allreduce_ms_alone"] * 1e-3) / 1e9
flat._hooks = [p.register_post_accumulate_grad_hook(flat._on_grad) for p in flat.params] if world > 1 else []
if world > 1:      # the un-synchronised variant above let the replicas drift apart: start the DP phase from rank 0's state
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, 0)
    for st in opt.state.values():
        for v in st.values():
            if torch.is_tensor(v) and v.is_cuda:
                dist.broadcast(v, 0)
    bbbp_b200.autograd.clear_weight_cache()
res["step_ms_bucketed_overlapped"] = timed(step_overlapped)
chk = torch.tensor([float(sum(p.double().sum() for p in model.parameters()))], device=dev, dtype=torch.float64)
if world > 1:
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    res["replicas_identical"] = bool(lo.item() == hi.item())
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
