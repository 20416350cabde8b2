"""Achieved HBM GB/s of the memory-bound kernels at sizes far beyond L2 (north_star: >= 70 % of HBM bandwidth on the
elementwise / norm kernels).  Bytes are algorithmic (each operand read or written once); peak = MEASURED_PEAKS.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
from bbbp_b200 import ops
dev = torch.device("cuda:0")
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6552.6
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
rows = []
def report(name, nbytes, ms):
    gbs = nbytes / ms / 1e6
    rows.append((name, nbytes / 1e6, ms, gbs, gbs / peak))
    print(f"{name:46s} {nbytes / 1e6:9.1f} MB {ms:8.3f} ms {gbs:8.0f} GB/s  {gbs / peak:5.2f} of {peak:.0f}", flush=True)

R = 1 << 21                                    # 2 M rows x 167 floats = 1.4 GB per operand
x, res = torch.randn(R, 167, device=dev), torch.randn(R, 167, device=dev)
g, b = torch.ones(167, device=dev), torch.zeros(167, device=dev)
report("add + LayerNorm fwd (x, res -> y)", 3 * x.numel() * 4, t(lambda: ops.add_layernorm_fwd(x, res, g, b)))
y, s_, mean, rstd, _ = ops.add_layernorm_fwd(x, res, g, b, save=True)
dy = torch.randn_like(x)
report("LayerNorm bwd (dy, s -> dx, dgamma, dbeta)", 3 * x.numel() * 4 + 2 * x.numel() * 4, t(lambda: ops.layernorm_bwd(dy, s_, mean, rstd, g)))
del y, s_, dy, res
report("cast fp32 -> bf16 (pitch 168)", x.numel() * 4 + R * 168 * 2, t(lambda: ops.cast16(x, 0)))
report("cast fp32 -> fp16 hi + lo pair (pitch 168)", x.numel() * 4 + 2 * R * 168 * 2, t(lambda: ops.cast16(x, 1, want_lo=True)))
x2 = torch.randn(R, 167, device=dev)
report("ReLU backward (dy, y -> dx)", 3 * x.numel() * 4, t(lambda: ops.act_bwd(x2, x, "relu")))
del x2
report("dropout (x -> y)", 2 * x.numel() * 4, t(lambda: ops.dropout(x, 0.1, 7)))
report("column sum (bias gradient)", x.numel() * 4, t(lambda: ops.colsum(x)))
del x
n = 16384
img8 = torch.randint(0, 256, (n, 3, 128, 128), device=dev, dtype=torch.uint8)
report("uint8 depiction -> z-scored fp32", img8.numel() * 5, t(lambda: ops.u8_zscore(img8.reshape(n, -1))))
report("uint8 depiction statistics", img8.numel(), t(lambda: ops.u8_image_stats(img8)))
white = torch.full((n, 3, 128, 128), 255, dtype=torch.uint8)
white[(torch.rand(n, 1, 128, 128) < 0.07).expand(-1, 3, -1, -1)] = 40
sd = bbbp_b200.SparseDepictions.encode(white.numpy(), pin=False)
m_, v_, o_ = sd.mask.to(dev), sd.values.to(dev), sd.offsets.to(dev)
report("sparse depictions -> uint8 CHW (7 % marked)", sd.nbytes() + white.numel(), t(lambda: ops.decode_sparse_depictions(m_, v_, o_)))
del white, m_, v_, o_
packed = torch.randint(0, 256, (1 << 22, 21), device=dev, dtype=torch.uint8)
report("packed bits -> z-scored fp32 (167 bits)", packed.numel() + (1 << 22) * 167 * 4, t(lambda: ops.unpack_zscore(packed, 167)))
del img8, packed
a = torch.randn(2048, 64, 64, 128, device=dev).bfloat16()
report("2x2 max-pool NHWC bf16", a.numel() * 2 * 1.25, t(lambda: ops.maxpool2x2_nhwc_bf16(a)))
a2 = torch.randn(512, 64, 64, 64, device=dev).bfloat16()
report("im2col 3x3 bf16 (C = 64)", a2.numel() * 2 * 10, t(lambda: ops.im2col3x3_bf16(a2)))
del a, a2
params = [torch.randn(1 << 24, device=dev, requires_grad=True) for _ in range(4)]
for p in params: p.grad = torch.randn_like(p)
opt = bbbp_b200.AdamW(params, lr=1e-4, weight_decay=1e-5)
report("fused AdamW (64 M parameters, 28 B each)", sum(p.numel() for p in params) * 28, t(opt.step))
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/hbm_kernels.txt", "w") as fh:
    fh.write(f"# python tools/hbm_kernels.py -- algorithmic bytes / CUDA-event time, peak {peak:.1f} GB/s (MEASURED_PEAKS.json, copy read+write)\n")
    for name, mb, ms, gbs, frac in rows:
        fh.write(f"{name:46s} {mb:9.1f} MB {ms:8.3f} ms {gbs:8.0f} GB/s  {frac:5.2f}\n")
