"""Fused feed-forward + LayerNorm kernel (bbbp_ffn_layernorm16) against the two-GEMM + LayerNorm route it replaces:
CUDA-event time per call at the encoder's shape (d = 167, hidden = 2048), rows = 2 048 ... 65 536."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
from bbbp_b200 import ops
dev = torch.device("cuda:0"); torch.manual_seed(0)
d, hidden, fmt = 167, 2048, int(os.environ.get("FMT", "0"))
ldq = 168
rows_list = [int(r) for r in os.environ.get("ROWS", "2048,8192,16384,65536").split(",")]
reps = int(os.environ.get("REPS", "20"))
w1, b1 = torch.randn(hidden, d, device=dev) * d ** -0.5, torch.randn(hidden, device=dev) * 0.1
w2, b2 = torch.randn(d, hidden, device=dev) * hidden ** -0.5, torch.randn(d, device=dev) * 0.1
g, be = torch.ones(d, device=dev), torch.zeros(d, device=dev)
w1_16, _ = ops.cast16(w1, fmt); w2_16, _ = ops.cast16(w2, fmt)
def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for rows in rows_list:
    x = torch.randn(rows, d, device=dev)
    x32 = torch.zeros(rows, ldq, device=dev); x32[:, :d] = x
    x16, _ = ops.cast16(x, fmt, ld=ldq)
    def fused():
        return ops.ffn_layernorm16(x16, d, w1_16, b1, w2_16, b2, x32, g, be, 1e-5, ld_y=ldq, ld16=ldq, fmt=fmt)
    def unfused():
        _, h16 = ops.gemm_bf16(x16, d, w1_16, hidden, bias=b1, act="relu", out_f32=False, out_bf16=True, fmt=fmt)
        f32, _ = ops.gemm_bf16(h16, hidden, w2_16, d, bias=b2, residual=x32, ld_out=ldq, fmt=fmt)
        return ops.layernorm_fwd_pitched(f32, d, g, be, 1e-5, ld_y=ldq, bf16_ld=ldq, fmt=fmt)
    tf, tu = timed(fused), timed(unfused)
    flop = 4.0 * rows * d * hidden
    print(f"rows {rows:6d}: fused {tf:8.1f} us ({flop / tf / 1e6:7.1f} TFLOP/s)   two GEMMs + LayerNorm {tu:8.1f} us ({flop / tu / 1e6:7.1f} TFLOP/s)")
