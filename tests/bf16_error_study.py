"""Distribution of |bf16 tensor-core output - fp32 oracle output| over a B3DB-sized set (diagnostic for the tolerance)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
from oracle import nets
torch.manual_seed(0)
for F in (167, 2048):
    torch.manual_seed(0)
    ref = nets.build("tcnn", F, 128).eval()
    ours = bbbp_b200.MixedInputModel(F, 128); ours.load_state_dict(ref.state_dict()); ours.cuda().eval()
    n, bs = (1058, 256) if F == 167 else (256, 64)
    g = torch.Generator().manual_seed(5)
    bits = (torch.rand(n, F, generator=g) < (0.25 if F == 167 else 0.022)).float(); bits[:, 0] = 0
    fp = (bits - bits.mean(1, keepdim=True)) / bits.std(1, unbiased=False, keepdim=True)
    img = torch.randn(n, 49152, generator=g)
    with torch.no_grad():
        want = torch.cat([ref(fp[i:i + bs], img[i:i + bs]).reshape(-1) for i in range(0, n, bs)])
        for prec in ("fp32", "bf16"):
            ours.set_precision(prec)
            got = ours.predict_batches(fp.cuda(), img.cuda(), bs).cpu()
            d = (got - want).abs()
            print(f"F={F} {prec}: max |d| {float(d.max()):.3e}  mean {float(d.mean()):.3e}  output std {float(want.std()):.3e} range [{float(want.min()):.3f}, {float(want.max()):.3f}]")
