"""How many cycles does one tcgen05.mma cost as a function of N in the SW128 K-major GEMM kernel? (diagnostic)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
from bbbp_b200 import ops
M, K = 74 * 128, int(os.environ.get('K', 4096))   # A = 77 MB bf16: L2-resident after the first pass
a16 = ops.cast_bf16(torch.randn(M, K, device="cuda"))
for N in (64, 128, 256):
    w16 = ops.cast_bf16(torch.randn(N, K, device="cuda") * 0.01)
    # force the tile width = N: N<=64 -> BN=64; N=128 -> BN=128; N=256: BN=128 x2 tiles unless N>=512 (BN=256)
    fn = lambda: ops.gemm_bf16(a16, K, w16, N, out_f32=False, out_bf16=True)
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tiles_n = 1 if N <= 128 else 2
    mmas_per_cta = (K // 16)
    print(f"N={N}: {ms * 1e3:.1f} us, {2 * M * N * K / ms / 1e9:.0f} TFLOP/s, ~{ms * 1e-3 * 1.9e9 / mmas_per_cta / tiles_n:.0f} cycles per MMA (BN={min(N,128)})")
