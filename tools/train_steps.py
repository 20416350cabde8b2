"""A few fp32 training steps of the canonical net (profiling driver for the backward kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200

B = int(os.environ.get("B", 32)); STEPS = int(os.environ.get("STEPS", 4))
dev = torch.device("cuda:0"); torch.manual_seed(0)
m = bbbp_b200.MixedInputModel(167, 128).to(dev); bbbp_b200.zero_dropout(m); m.train()
opt = bbbp_b200.AdamW(m.parameters(), lr=1e-4, weight_decay=1e-5); crit = bbbp_b200.MSELoss()
fp, img, y = torch.randn(B, 167, device=dev), torch.randn(B, 49152, device=dev), torch.randn(B, device=dev)
for i in range(STEPS):
    if i == STEPS - 1:
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    opt.zero_grad(); loss = crit(m(fp, img).squeeze(), y); loss.backward(); opt.step()
e1.record(); torch.cuda.synchronize(); print("last step ms", e0.elapsed_time(e1), "loss", float(loss.detach()))
