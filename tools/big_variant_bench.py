"""Inference rate of the big variant (20250107_network.py: 12-layer encoder, conv stack 3 -> 64 -> 128 -> 256) in the bf16
tensor-core mode: the implicit-GEMM convolutions (BBBP_IMPLICIT_CONV=1, default) against the explicit im2col route (=0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
dev = torch.device("cuda:0"); torch.manual_seed(0)
m = bbbp_b200.build("tcnn_big", 167, 128).to(dev).eval().set_precision("bf16")
m.use_cuda_graphs = False        # (a captured graph would replay whichever route was captured first)
n = int(os.environ.get("N", 1024))
fp, img = torch.randn(n, 167, device=dev), torch.randn(n, 49152, device=dev)
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
outs = {}
with torch.no_grad():
    for implicit in (False, True):
        m.implicit_conv = implicit
        ms = t(lambda: m.predict_batches(fp, img, 256, max_rows_per_pass=n))
        outs[implicit] = m.predict_batches(fp, img, 256, max_rows_per_pass=n)
        print(f"big variant bf16, {n} molecules, implicit_conv={implicit}: {ms:8.2f} ms = {n / ms * 1e3:9.0f} mol/s", flush=True)
print("max |implicit - im2col| =", float((outs[True] - outs[False]).abs().max()), "spread", float(outs[False].std()))
