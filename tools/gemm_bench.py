"""CUDA-event timing of the tcgen05 GEMM on the encoder's shapes (diagnostic, not a bench line)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bbbp_b200
from bbbp_b200 import ops

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

M = int(os.environ.get("M", 8192))
for (N, K, res, o32, o16, act, ldo) in [(504, 167, False, False, True, None, None), (167, 167, True, True, False, None, None),
                                         (167, 167, True, True, False, None, 168), (167, 167, False, False, True, None, None),
                                         (2048, 167, False, False, True, "relu", None), (2048, 167, False, True, False, "relu", None),
                                         (167, 2048, True, True, False, None, None), (167, 2048, False, False, True, None, None),
                                         (128, 167, False, True, False, "relu", None), (128, 256, False, True, False, "relu", None),
                                         (2048, 2048, False, False, True, None, None), (4096, 4096, False, False, True, None, None)]:
    a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") * 0.05; b = torch.randn(N, device="cuda")
    a16, w16 = ops.cast_bf16(a), ops.cast_bf16(w)
    r = torch.randn(M, N if ldo is None else ldo, device="cuda") if res else None
    fn = lambda: ops.gemm_bf16(a16, K, w16, N, bias=b, residual=r, act=act, out_f32=o32, out_bf16=o16, ld_out=ldo)
    us = t(fn)
    print(f"M={M} N={N:5d} K={K:5d} res={int(res)} f32={int(o32)} bf16={int(o16)} ld={ldo}: {us:8.1f} us  {2*M*N*K/us/1e6:8.1f} TFLOP/s")
