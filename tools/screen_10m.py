"""BASELINE configs[3]: virtual-screening inference over a synthetic ZINC-shaped library of 10 M molecules, sharded by
whole reference batches over the ranks (python -m torch.distributed.run ... tools/screen_10m.py), one NCCL gather of the
scores.  MODEL=morgan (default, SURVEY cfg4: 2048-bit fingerprints packed to 256 B, 160 M-parameter network) or maccs
(167 bits in 21 B); PRECISION=strict (default) | fp16 | bf16.  Molecules are packed bit rows + uint8 3x128x128 depictions generated ON THE DEVICE shard by shard from a
generator keyed by the GLOBAL shard index (shards are cut on a global grid, so any rank count sees the same library and the
score checksum must not depend on it); generation is outside the timer, unpack + z-score + in-kernel image normalisation + forward are inside (CUDA events per shard, max over ranks)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bbbp_b200

N = int(os.environ.get("N", 10_000_000)); BATCH = 256; SHARD = int(os.environ.get("SHARD", 16384))
MODEL = os.environ.get("MODEL", "morgan"); PRECISION = os.environ.get("PRECISION", "strict")
F_BITS = 2048 if MODEL == "morgan" else 167
PACKED = (F_BITS + 7) // 8
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = bbbp_b200.MixedInputModel(F_BITS, 128).to(dev).eval().set_precision(PRECISION)
a, b = bbbp_b200.partition_batches(N, BATCH, world, rank)
scores = torch.empty(b - a, device=dev, dtype=torch.float32)

def shard_data(start, stop):
    """Global shard [start, stop): start is a multiple of SHARD."""
    g = torch.Generator(device=dev).manual_seed(20250113 + start // SHARD)
    n = stop - start
    if MODEL == "morgan":       # ~45 on-bits of 2048 (Morgan radius 2): AND of five uniform bytes -> bit density 1/32
        packed = torch.randint(0, 256, (n, PACKED), generator=g, device=dev, dtype=torch.uint8)
        for _ in range(4):
            packed &= torch.randint(0, 256, (n, PACKED), generator=g, device=dev, dtype=torch.uint8)
    else:
        packed = torch.randint(0, 256, (n, PACKED), generator=g, device=dev, dtype=torch.uint8)
    strokes = torch.rand((n, 1, 128, 128), generator=g, device=dev) < 0.06        # dark strokes on a white depiction
    img = torch.where(strokes, 40, 255).to(torch.uint8).expand(-1, 3, -1, -1).contiguous()
    return packed, img

with torch.no_grad():
    p, i = shard_data(0, SHARD)
    grid = list(range(a // SHARD * SHARD, b, SHARD))
    sizes = {min(b, g0 + SHARD) - max(a, g0) for g0 in grid}              # full shards + this rank's ragged first / last one
    for rows in sorted(sizes):                                            # untimed: allocator pools, TMA maps, tail-batch graphs
        for _ in range(2):
            model.predict_batches_packed(p[:rows], i[:rows], BATCH, max_rows_per_pass=SHARD)
    if world > 1:
        bbbp_b200.gather_scores(torch.zeros(b - a, device=dev), N, BATCH)       # NCCL communicator set-up, untimed
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall0 = time.perf_counter(); dev_ms = 0.0; shard_ms = []
    for g0 in grid:
        p, i = shard_data(g0, min(N, g0 + SHARD))
        lo, hi = max(a, g0), min(b, g0 + SHARD)                                  # this rank's rows of the global shard
        p, i = p[lo - g0:hi - g0], i[lo - g0:hi - g0]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        scores[lo - a:hi - a].copy_(model.predict_batches_packed(p, i, BATCH, max_rows_per_pass=SHARD))
        e1.record(); e1.synchronize()
        dev_ms += e0.elapsed_time(e1)
        shard_ms.append(round(e0.elapsed_time(e1), 2))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    full = bbbp_b200.gather_scores(scores, N, BATCH)
    e1.record(); e1.synchronize()
    gather_ms = e0.elapsed_time(e1)
    t = torch.tensor([dev_ms + gather_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall = time.perf_counter() - wall0
if rank == 0:
    print(json.dumps({"workload": f"screen {N} synthetic ZINC-shaped molecules (packed {F_BITS}-bit fingerprints + uint8 depictions), "
                                  f"batch 256, {MODEL} network, {PRECISION} mode",
                      "n_molecules": N, "n_gpus": world, "device_ms_max_over_ranks": float(t), "gather_ms": gather_ms,
                      "molecules_per_s": N / (float(t) * 1e-3), "wall_s_including_generation": wall,
                      "score_checksum": float(full.double().sum()), "finite": bool(torch.isfinite(full).all()),
                      "shard_ms_first_last": shard_ms[:4] + shard_ms[-3:], "shard_ms_median": sorted(shard_ms)[len(shard_ms) // 2]}))
if world > 1:
    dist.destroy_process_group()
