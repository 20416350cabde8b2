"""CUDA-event timing + cycle probe of the strict-mode conv kernels: (hi, lo) pairs vs background-referenced (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
from bbbp_b200 import ops
from bbbp_b200._lib import lib, check
N = int(os.environ.get("N", 16384))
def t(fn, n=6):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
torch.manual_seed(0)
W1, b1 = torch.randn(32, 3, 3, 3, device="cuda") * 0.2, torch.randn(32, device="cuda") * 0.1
W2, b2 = torch.randn(64, 32, 3, 3, device="cuda") * 0.05, torch.randn(64, device="cuda") * 0.1
w1, w2 = ops.conv3x3_prepare_bf16(W1, 1), ops.conv3x3_prepare_bf16(W2, 1)
img8 = torch.full((N, 3, 128, 128), 255, dtype=torch.uint8, device="cuda")
img8[(torch.rand(N, 1, 128, 128, device="cuda") < 0.06).expand(-1, 3, -1, -1)] = 40
stats = ops.u8_image_stats(img8)
img = ops.u8_zscore(img8.view(N, -1)).view(N, 3, 128, 128)
bg1 = ops.image_background(img)
ws1, ws2 = ops.fc_weight_channel_sums(W1, 3, 9), ops.fc_weight_channel_sums(W2, 32, 9)
tab1, neg2 = ops.bg_layer(ws1, b1, bg1, fmt=1, want_neg16=True)
tab2, _ = ops.bg_layer(ws2, b2, tab1[:, 1], fmt=-1)
y1, y1lo = ops.conv1_from_image_bf16(img, w1, b1, fmt=1, split=True)
y1b = ops.conv1_from_image_bg(img, w1, None, bg1, tab1, fmt=1, split=True)
bg1_u8 = ops.image_background(img8, stats)          # carries the raw background bytes the exact-integer form reads
cases = [
    ("conv1 fp16 one pass (fp32 planes)", lambda: ops.conv1_from_image_bf16(img, w1, b1, fmt=1)),
    ("conv1 strict pairs  (fp32 planes)", lambda: ops.conv1_from_image_bf16(img, w1, b1, fmt=1, split=True)),
    ("conv1 BG 2 passes   (fp32 planes)", lambda: ops.conv1_from_image_bg(img, w1, None, bg1, tab1, fmt=1, split=True)),
    ("conv1 BG 1 pass     (fp32 planes)", lambda: ops.conv1_from_image_bg(img, w1, None, bg1, tab1, fmt=1, split=False)),
    ("conv1 fp16 one pass (uint8)", lambda: ops.conv1_from_image_bf16(img8, w1, b1, stats, fmt=1)),
    ("conv1 BG 2 passes   (uint8)", lambda: ops.conv1_from_image_bg(img8, w1, stats, bg1, tab1, fmt=1, split=True)),
    ("conv1 BG exact ints, 1 pass (uint8)", lambda: ops.conv1_from_image_bg(img8, w1, stats, bg1_u8, tab1, fmt=1, split=False)),
    ("conv2 fp16 one pass", lambda: ops.conv3x3_relu_pool_bf16(y1, w2, b2, 64, fmt=1)),
    ("conv2 strict pairs", lambda: ops.conv3x3_relu_pool_bf16(y1, w2, b2, 64, fmt=1, x_lo=y1lo)),
    ("conv2 BG", lambda: ops.conv3x3_relu_pool_bg(y1b, w2, neg2, tab2, 64, fmt=1)),
    ("image_background", lambda: ops.image_background(img)),
    ("bg_layer 3->32", lambda: ops.bg_layer(ws1, b1, bg1, fmt=1, want_neg16=True)),
    ("bg_layer 32->64", lambda: ops.bg_layer(ws2, b2, tab1[:, 1], fmt=-1)),
]
print(f"N={N}")
for label, fn in cases:
    print(f"{label:36s} {t(fn):.3f} ms")
names = ["mma: wait acc_empty", "mma: wait full", "mma: issue+commit", "mma: tiles", "epi: wait acc_full", "epi: tmem+math+sts",
         "epi: wait buffer free (bar1)", "epi: fence + arrive", "prod: wait empty", "prod: issue loads", "prod: wait data+arrive"]
for label, fn in cases[:10]:
    probe = torch.zeros(16, dtype=torch.int64, device="cuda")
    check(lib.bbbp_debug_conv_probe(probe.data_ptr()))
    fn(); torch.cuda.synchronize()
    check(lib.bbbp_debug_conv_probe(None))
    p = probe.cpu().tolist(); tiles = max(1, p[3])
    print(f"--- {label}: {tiles} tiles on CTA 0; cycles per tile: " + ", ".join(f"{nm} {p[k] / tiles:.0f}" for k, nm in enumerate(names) if k != 3))
