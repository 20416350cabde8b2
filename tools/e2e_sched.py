"""predict_from_host with explicit chunk schedules (monkeypatched _pipeline_spans)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
dev = torch.device("cuda:0"); torch.manual_seed(0)
m = bbbp_b200.MixedInputModel(167, 128).to(dev).eval().set_precision("bf16")
n = int(os.environ.get('N', 16384))
packed = torch.randint(0, 256, (n, 21), dtype=torch.uint8).pin_memory()
img8 = torch.randint(0, 256, (n, 3, 128, 128), dtype=torch.uint8).pin_memory()
out_host = torch.empty(n, dtype=torch.float32).pin_memory()
scheds = [[1024] * (n // 1024), [2048] * (n // 2048), [1536] * (n // 1536) + ([n % 1536] if n % 1536 else []), [768] * (n // 768) + ([n % 768] if n % 768 else []),
          [4096] * (n // 4096)]
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
