"""Per-step wall time of the overlapped host pipeline (diagnostic for run-to-run variance)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
dev = torch.device("cuda:0"); torch.manual_seed(0)
m = bbbp_b200.MixedInputModel(167, 128).to(dev).eval().set_precision("bf16")
n = 8192
packed = torch.randint(0, 256, (n, 21), dtype=torch.uint8).pin_memory()
img8 = torch.randint(0, 256, (n, 3, 128, 128), dtype=torch.uint8).pin_memory()
out = torch.empty(n, dtype=torch.float32).pin_memory()
chunk = int(os.environ.get("CHUNK", 2048))
for _ in range(3):
    m.predict_from_host(packed, img8, 256, chunk_molecules=chunk, packed=True, out_host=out)
torch.cuda.synchronize()
ts = []
for i in range(30):
    t0 = time.perf_counter()
    m.predict_from_host(packed, img8, 256, chunk_molecules=chunk, packed=True, out_host=out)
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print("chunk", chunk, "per-step ms:", " ".join(f"{t:.1f}" for t in ts))
# raw H2D bandwidth of the same buffers
d = torch.empty_like(img8, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): d.copy_(img8, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"raw H2D {img8.numel() / dt / 1e9:.1f} GB/s")

# ---- where do the spikes come from? allocator activity and resident-input timing of the same chunks ----------------
import statistics
st0 = torch.cuda.memory_stats()
ts = []
for i in range(20):
    t0 = time.perf_counter()
    m.predict_from_host(packed, img8, 256, chunk_molecules=chunk, packed=True, out_host=out)
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
st1 = torch.cuda.memory_stats()
for k in ("num_device_alloc", "num_device_free", "num_alloc_retries", "allocation.all.allocated"):
    print(k, st1.get(k, 0) - st0.get(k, 0))
pk_d, im_d = packed.to(dev), img8.to(dev)
ts2 = []
for i in range(20):
    t0 = time.perf_counter()
    for a in range(0, n, chunk):
        m.predict_batches_packed(pk_d[a:a + chunk], im_d[a:a + chunk], 256, max_rows_per_pass=chunk)
    torch.cuda.synchronize()
    ts2.append((time.perf_counter() - t0) * 1e3)
print("resident, same chunking, per-step ms:", " ".join(f"{t:.1f}" for t in ts2))
ts3 = []
for i in range(20):
    t0 = time.perf_counter()
    a_ = packed.to(dev, non_blocking=True); b_ = img8.to(dev, non_blocking=True)
    torch.cuda.synchronize()
    ts3.append((time.perf_counter() - t0) * 1e3)
print("H2D only per-step ms:", " ".join(f"{t:.1f}" for t in ts3))
