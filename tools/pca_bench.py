"""P16: sklearn PCA.transform of the depiction features, (N, 49152) -> (N, 128), on the tensor cores (strict: centring fused
into the fp32 -> (hi, lo) fp16 split, both operands split, split-K tcgen05 GEMM) vs the CUDA-core fp32 GEMM; and the chunked
per-feature standardisation (N3) on the same matrix.  CUDA events, inputs larger than L2."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
dev = torch.device("cuda:0"); torch.manual_seed(0)
N, D, k = int(os.environ.get("N", 8192)), 49152, 128
x = torch.randn(N, D, device=dev); comp = torch.randn(k, D, device=dev) / 200; mu = x.mean(0)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
res = {"rows": N, "features": D, "components": k}
want = ((x[:64].double() - mu.double()) @ comp.double().T)
for mode in ("strict", "fp16", "bf16", "fp32"):
    ms = t(lambda: bbbp_b200.pca_transform(x, mu, comp, precision=mode))
    err = float((bbbp_b200.pca_transform(x[:64].contiguous(), mu, comp, precision=mode).double() - want).abs().max())
    res[mode] = {"ms": ms, "rows_per_s": N / ms * 1e3, "GB_per_s_of_input": N * D * 4 / ms / 1e6, "tflops_algorithmic": 2 * N * D * k / ms / 1e9,
                 "max_abs_err_vs_float64": err}
    print(mode, res[mode], flush=True)
ms = t(lambda: bbbp_b200.standardize_chunks(x, 100))
res["standardize_chunks_100"] = {"ms": ms, "GB_per_s_algorithmic": 2 * N * D * 4 / ms / 1e6}
print(json.dumps(res))
