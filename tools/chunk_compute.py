"""GPU time of one staged chunk (graph replay) vs its H2D copy, per chunk size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
dev = torch.device("cuda:0"); torch.manual_seed(0)
m = bbbp_b200.MixedInputModel(167, 128).to(dev).eval().set_precision("bf16")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
with torch.no_grad():
    for rows in (256, 512, 1024, 2048, 4096):
        packed_h = torch.randint(0, 256, (rows, 21), dtype=torch.uint8).pin_memory()
        img_h = torch.randint(0, 256, (rows, 3, 128, 128), dtype=torch.uint8).pin_memory()
        packed, img = packed_h.to(dev), img_h.to(dev)
        g = t(lambda: m._score_staged_chunk(0, packed, img, 256, rows, True))
        m.use_cuda_graphs = False
        e = t(lambda: m.predict_batches_packed(packed, img, 256, max_rows_per_pass=rows))
        m.use_cuda_graphs = True
        c = t(lambda: (packed.copy_(packed_h, non_blocking=True), img.copy_(img_h, non_blocking=True)))
        print(f"rows {rows}: graph {g:.3f} ms  eager {e:.3f} ms  h2d {c:.3f} ms", flush=True)
