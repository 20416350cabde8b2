"""Per-node latency of dependent kernels inside a CUDA graph: skinny GEMM alone vs alternating with other small kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
from bbbp_b200 import ops
dev = torch.device("cuda:0")
x = torch.randn(32, 167, device=dev); w = torch.randn(167, 167, device=dev) * 0.05
g_, b_ = torch.ones(167, device=dev), torch.zeros(167, device=dev)
w2 = torch.randn(2048, 167, device=dev) * 0.05; w3 = torch.randn(167, 2048, device=dev) * 0.02

def chain(kind, n=60):
    y = x
    for i in range(n):
        if kind == "gemm":
            y = ops.gemm_f32(y, w, trans_b=True, split_k=0)
        elif kind == "gemm+ln":
            y = ops.gemm_f32(y, w, trans_b=True, split_k=0)
            y = ops.add_layernorm_fwd(y, None, g_, b_)[0]
        elif kind == "ln":
            y = ops.add_layernorm_fwd(y, None, g_, b_)[0]
        elif kind == "copy":
            y = ops.copy2d(y, torch.empty_like(y))
        elif kind == "ffn":
            h = ops.gemm_f32(y, w2, trans_b=True, act="relu", split_k=0)
            y = ops.gemm_f32(h, w3, trans_b=True, split_k=0)
        elif kind == "gemm+copy":
            y = ops.gemm_f32(y, w, trans_b=True, split_k=0)
            y = ops.copy2d(y, torch.empty_like(y))
    return y

for kind, per in (("gemm", 1), ("ln", 1), ("copy", 1), ("gemm+ln", 2), ("gemm+copy", 2), ("ffn", 3)):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        chain(kind, 3)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        chain(kind)
    for _ in range(3): g.replay()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 / 60 * 1e3
    print(f"{kind:10s}: {us:6.2f} us per iteration ({us / per:5.2f} us per kernel node)", flush=True)
