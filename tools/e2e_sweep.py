"""predict_from_host on the compact formats: chunk-size sweep (pipeline fill/drain vs per-chunk kernel efficiency)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
dev = torch.device("cuda:0"); torch.manual_seed(0)
m = bbbp_b200.MixedInputModel(167, 128).to(dev).eval().set_precision(os.environ.get("PRECISION", "bf16"))
n = int(os.environ.get("N", 8192))
packed = torch.randint(0, 256, (n, 21), dtype=torch.uint8).pin_memory()
img8 = torch.randint(0, 256, (n, 3, 128, 128), dtype=torch.uint8).pin_memory()
out_host = torch.empty(n, dtype=torch.float32).pin_memory()
if os.environ.get("SPARSE", "0") == "1":            # mostly-white depictions in the lossless sparse encoding
    img8 = torch.full((n, 3, 128, 128), 255, dtype=torch.uint8)
    strokes = torch.rand(n, 1, 128, 128) < 0.06
    img8[strokes.expand(-1, 3, -1, -1)] = 40
    img8 = bbbp_b200.SparseDepictions.encode(img8.numpy())
res = {}
schedules = [(1024, 5120, 10240), (2048, 14336), (2048, 6144, 8192), (1024, 3072, 4096, 8192), (4096, 12288), (2048, 4096, 10240)]
chunks = [256, 512, 1024, 2048, 4096, 8192, 16384] if os.environ.get("SCHEDULES", "0") != "only" else []
for chunk in chunks + (schedules if os.environ.get("SCHEDULES", "0") != "0" else []):
    if (max(chunk) if isinstance(chunk, tuple) else chunk) > n: continue
    fn = lambda: m.predict_from_host(packed, img8, 256, chunk_molecules=chunk, packed=True, out_host=out_host)
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8): fn()
    e1.record(); torch.cuda.synchronize()
    res[str(chunk)] = e0.elapsed_time(e1) / 8
    print(chunk, res[str(chunk)], n / res[str(chunk)] * 1e3, flush=True)
# H2D alone
if not torch.is_tensor(img8): sys.exit(0)
d = torch.empty_like(img8, device=dev)
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(8): d.copy_(img8, non_blocking=True)
e1.record(); torch.cuda.synchronize()
res["h2d_only_ms"] = e0.elapsed_time(e1) / 8
print(json.dumps(res))
