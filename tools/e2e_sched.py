"""predict_from_host with explicit chunk schedules (monkeypatched _pipeline_spans)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
dev = torch.device("cuda:0"); torch.manual_seed(0)
m = bbbp_b200.MixedInputModel(167, 128).to(dev).eval().set_precision("bf16")
n = 8192
packed = torch.randint(0, 256, (n, 21), dtype=torch.uint8).pin_memory()
img8 = torch.randint(0, 256, (n, 3, 128, 128), dtype=torch.uint8).pin_memory()
out_host = torch.empty(n, dtype=torch.float32).pin_memory()
scheds = [[2048] * 4, [4096, 4096], [1024] * 8, [1024, 2048, 2048, 2048, 1024], [2048, 4096, 2048], [3072, 3072, 2048], [4096, 3072, 1024],
          [4096, 2048, 1024, 1024], [1024, 3072, 3072, 1024], [512, 1536, 2048, 2048, 1536, 512], [2048, 2048, 2048, 1024, 1024], [3072, 3072, 1024, 1024],
          [1024, 3072, 2048, 1024, 1024], [4096, 2048, 2048]]
for sizes in scheds:
    spans, a = [], 0
    for s in sizes: spans.append((a, a + s)); a += s
    type(m)._pipeline_spans = staticmethod(lambda n_, c_, b_, spans=spans: spans)
    fn = lambda: m.predict_from_host(packed, img8, 256, chunk_molecules=max(sizes), packed=True, out_host=out_host)
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 8
    print(f"{ms:7.3f} ms {n / ms * 1e3:10.0f} mol/s  {sizes}", flush=True)
