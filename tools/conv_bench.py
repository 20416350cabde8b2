"""CUDA-event timing of the tcgen05 conv kernels (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
from bbbp_b200 import ops
N = int(os.environ.get("N", 8192))
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
w1 = ops.conv3x3_prepare_bf16(torch.randn(32, 3, 3, 3, device="cuda")); b1 = torch.randn(32, device="cuda")
w2 = ops.conv3x3_prepare_bf16(torch.randn(64, 32, 3, 3, device="cuda") * 0.05); b2 = torch.randn(64, device="cuda")
img = torch.randn(N, 3, 128, 128, device="cuda")
img8 = torch.randint(0, 256, (N, 3, 128, 128), device="cuda", dtype=torch.uint8)
stats = ops.u8_image_stats(img8)
x8 = ops.image_to_nhwc8_bf16(img)
y1 = ops.conv1_from_image_bf16(img, w1, b1)
print(f"N={N}")
print(f"conv1 from fp32 planes : {t(lambda: ops.conv1_from_image_bf16(img, w1, b1)):.3f} ms")
print(f"conv1 from uint8 planes: {t(lambda: ops.conv1_from_image_bf16(img8, w1, b1, stats)):.3f} ms")
print(f"conv1 from NHWC8 bf16  : {t(lambda: ops.conv3x3_relu_pool_bf16(x8, w1, b1, 32)):.3f} ms")
print(f"u8 stats               : {t(lambda: ops.u8_image_stats(img8)):.3f} ms")
ms = t(lambda: ops.conv3x3_relu_pool_bf16(y1, w2, b2, 64))
print(f"conv2                  : {ms:.3f} ms  {N * 2 * 64 * 64 * 64 * 288 / ms / 1e9:.0f} TFLOP/s")

# ---- cycle probe of CTA 0's pipeline roles -------------------------------------------------------------------------
from bbbp_b200._lib import lib, check
names = ["mma: wait acc_empty", "mma: wait full", "mma: issue+commit", "mma: tiles", "epi: wait acc_full", "epi: tmem+math+sts",
         "epi: wait buffer free (bar1)", "epi: fence + arrive", "prod: wait empty", "prod: issue loads", "prod: wait data+arrive"]
for label, fn in [("conv2", lambda: ops.conv3x3_relu_pool_bf16(y1, w2, b2, 64)),
                  ("conv1 NHWC8", lambda: ops.conv3x3_relu_pool_bf16(x8, w1, b1, 32)),
                  ("conv1 fp32 planes", lambda: ops.conv1_from_image_bf16(img, w1, b1)),
                  ("conv1 uint8 planes", lambda: ops.conv1_from_image_bf16(img8, w1, b1, stats))]:
    probe = torch.zeros(16, dtype=torch.int64, device="cuda")
    check(lib.bbbp_debug_conv_probe(probe.data_ptr()))
    fn(); torch.cuda.synchronize()
    check(lib.bbbp_debug_conv_probe(None))
    p = probe.cpu().tolist(); tiles = max(1, p[3])
    print(f"--- {label}: {tiles} tiles on CTA 0; cycles per tile:")
    for k, nm in enumerate(names):
        if k != 3: print(f"    {nm:26s} {p[k] / tiles:9.0f}")
