import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
dev = torch.device("cuda:0"); torch.manual_seed(0)
m = bbbp_b200.MixedInputModel(167, 128).to(dev).eval().set_precision("bf16")
n = 8192
packed = torch.randint(0, 256, (n, 21), dtype=torch.uint8).pin_memory()
img8 = torch.randint(0, 256, (n, 3, 128, 128), dtype=torch.uint8).pin_memory()
out = torch.empty(n, dtype=torch.float32).pin_memory()
for i in range(4):
    m.predict_from_host(packed, img8, 256, chunk_molecules=2048, packed=True, out_host=out)
    torch.cuda.synchronize()
    p = getattr(m, "_host_pipe", None)
    print("step", i, "pipe", None if p is None else (p[0], p[1]), "reserved MiB", torch.cuda.memory_reserved() // 2**20, flush=True)
    snap = torch.cuda.memory_snapshot()
    big = sorted(((s["total_size"], s["stream"], sum(b["size"] for b in s["blocks"] if b["state"].startswith("active"))) for s in snap), reverse=True)[:8]
    print("   largest segments (size, stream, active bytes):", big)
