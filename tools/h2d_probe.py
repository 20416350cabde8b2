"""Per-GPU pinned-host -> device bandwidth with all ranks copying at once (explains the e2e scaling curve: the host side of
the box, not a kernel or a collective, bounds the host-fed screening rate).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 --master-port P tools/h2d_probe.py

Every rank copies TOTAL_MB from its own pinned buffer in chunks of CHUNK_MB on 1, 2 and 4 copy streams, all ranks starting
together; CUDA events per rank, rank 0 prints per-rank GB/s and the aggregate."""
import json, os, sys
import torch, torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
TOTAL = int(os.environ.get("TOTAL_MB", 1536)) << 20
host = torch.empty(TOTAL, dtype=torch.uint8).pin_memory(); host.fill_(1)
devbuf = torch.empty(TOTAL, dtype=torch.uint8, device=dev)
out = {"n_gpus": world, "total_MB_per_rank": TOTAL >> 20, "cases": []}
for chunk_mb in (1, 8, 48, 256):
    for streams in (1, 2, 4):
        chunk = chunk_mb << 20
        ss = [torch.cuda.Stream(dev) for _ in range(streams)]
        def run():
            for i, off in enumerate(range(0, TOTAL, chunk)):
                with torch.cuda.stream(ss[i % streams]):
                    devbuf[off:off + chunk].copy_(host[off:off + chunk], non_blocking=True)
        run(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        e0.record()
        for s in ss:
            s.wait_stream(cur)
        run()
        for s in ss:
            cur.wait_stream(s)
        e1.record(); torch.cuda.synchronize()
        gbs = torch.tensor([TOTAL / (e0.elapsed_time(e1) * 1e-3) / 1e9], device=dev)
        allg = [torch.zeros_like(gbs) for _ in range(world)]
        if world > 1:
            dist.all_gather(allg, gbs)
        else:
            allg = [gbs]
        out["cases"].append({"chunk_MB": chunk_mb, "streams": streams, "per_rank_GBs": [round(float(t), 1) for t in allg],
                             "aggregate_GBs": round(sum(float(t) for t in allg), 1)})
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
