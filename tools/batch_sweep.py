"""BASELINE.json configs[4]: inference batch-size sweep B = 2^0 .. 2^16, ONE forward call per B (the whole batch is one
attention scope, exactly as the reference would evaluate it).  Prints a table and writes gpurun_out/batch_sweep.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bbbp_b200

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = bbbp_b200.MixedInputModel(167, 128).to(dev).eval().set_precision(os.environ.get("PRECISION", "bf16"))
model.tensor_core_chunk = 8192
max_log2 = int(os.environ.get("MAX_LOG2", 16))
rows = []
g = torch.Generator(device=dev).manual_seed(1)
with torch.no_grad():
    for k in range(max_log2 + 1):
        B = 1 << k
        fp = torch.randn(B, 167, device=dev, generator=g)
        img = torch.randn(B, 3 * 128 * 128, device=dev, generator=g)
        reps = 20 if B <= 4096 else (5 if B <= 16384 else 2)
        for _ in range(3 if B <= 4096 else 1):
            model(fp, img)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = model(fp, img)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        flop = B * (206.0e6 + 2 * 6 * 2 * B * 167)           # + the S = B attention term
        rows.append({"batch": B, "latency_ms": ms, "molecules_per_s": B / ms * 1e3, "tflops": flop / ms / 1e9,
                     "attention": "QK^T GEMM with softmax epilogue + P V GEMM" if B < model.flash_min_seq else
                     "streaming-softmax tcgen05 kernel with out_proj + norm1 in its tail (logits stay on chip)"})
        print(f"B={B:6d}  {ms:10.3f} ms  {B / ms * 1e3:12.0f} mol/s  {flop / ms / 1e9:8.1f} TFLOP/s  {rows[-1]['attention']}", flush=True)
        del fp, img, out
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"model": "MixedInputModel(167,128)", "precision": model.precision, "rows": rows}, open("gpurun_out/batch_sweep.json", "w"), indent=1)
