import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
dev = torch.device("cuda:0"); torch.manual_seed(0)
m = bbbp_b200.MixedInputModel(167, 128).to(dev).eval().set_precision("bf16")
n = 8192
packed = torch.randint(0, 256, (n, 21), dtype=torch.uint8).pin_memory()
img8 = torch.randint(0, 256, (n, 3, 128, 128), dtype=torch.uint8).pin_memory()
out = torch.empty(n, dtype=torch.float32).pin_memory()
for _ in range(3):
    m.predict_from_host(packed, img8, 256, chunk_molecules=2048, packed=True, out_host=out)
torch.cuda.synchronize()
def segs():
    c = collections.Counter()
    for s in torch.cuda.memory_snapshot():
        c[(s["stream"], s["total_size"])] += 1
    return c
s0 = segs(); r0 = torch.cuda.memory_reserved()
for i in range(10):
    t0 = time.perf_counter()
    m.predict_from_host(packed, img8, 256, chunk_molecules=2048, packed=True, out_host=out)
    torch.cuda.synchronize()
    print(f"step {i}: {(time.perf_counter() - t0) * 1e3:.1f} ms reserved {torch.cuda.memory_reserved() / 1e6:.0f} MB  device_allocs {torch.cuda.memory_stats()['num_device_alloc']}")
s1 = segs()
print("new segments (stream, size) x count:", {k: v - s0.get(k, 0) for k, v in s1.items() if v != s0.get(k, 0)})
