"""Short workload for `ncu --set full`: one strict-mode pass over 4 096 molecules in reference batches of 256 (both tcgen05
convolutions, the GEMM template with its split passes, LayerNorm) and one pass with a 4 096-molecule attention scope (the
streaming-softmax kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
dev = torch.device("cuda:0"); torch.manual_seed(0)
m = bbbp_b200.MixedInputModel(167, 128).to(dev).eval().set_precision(os.environ.get("PRECISION", "strict"))
m.use_cuda_graphs = False
n = 4096
fp, img = torch.randn(n, 167, device=dev), torch.randn(n, 49152, device=dev)
img8 = torch.full((n, 3, 128, 128), 255, dtype=torch.uint8, device=dev)
img8[(torch.rand(n, 1, 128, 128, device=dev) < 0.06).expand(-1, 3, -1, -1)] = 40
bits = torch.randint(0, 256, (n, 21), dtype=torch.uint8, device=dev)
with torch.no_grad():
    for _ in range(2):
        m.predict_batches(fp, img, 256, max_rows_per_pass=n)      # 16 reference batches, fp32 contract
        m.predict_batches_packed(bits, img8, 256, max_rows_per_pass=n)   # packed bits + raw uint8 depictions (exact-integer conv1)
        m(fp, img)                                                # one 4 096-wide attention scope
torch.cuda.synchronize()
print("ok")
