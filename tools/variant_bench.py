"""Inference timing of the other variants (Morgan-2048 canonical net, big 20250107 net, MLP family) -- diagnostic."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
dev = torch.device("cuda:0")
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
for name, variant, F, img_dim, groups in [("Morgan-2048 canonical", "tcnn", 2048, 49152, 8), ("MACCS big (20250107)", "tcnn_big", 167, 49152, 4),
                                           ("MACCS no-fusion", "tcnn_nofusion", 167, 49152, 8), ("MLP opt (64,128)", "mlp", 64, 128, 256)]:
    torch.manual_seed(0)
    m = bbbp_b200.build(variant, F, 128).to(dev).eval()
    n = groups * 256
    fp, img = torch.randn(n, F, device=dev), torch.randn(n, img_dim, device=dev)
    for prec in ("fp32", "bf16"):
        m.set_precision(prec)
        with torch.no_grad():
            if variant.startswith("tcnn"):
                ms = t(lambda: m.predict_batches(fp, img, 256, max_rows_per_pass=n))
            else:
                ms = t(lambda: m(fp, img))
        print(f"{name:24s} {prec}: {n} molecules in {ms:8.2f} ms = {n / ms * 1e3:10.0f} mol/s", flush=True)
    del m, fp, img
# training step of the Morgan-2048 variant (BASELINE configs[2]: BCE loss, batch 32): eager loop body vs graph replay

torch.manual_seed(0)
m = bbbp_b200.build("tcnn", 2048, 128).to(dev); bbbp_b200.zero_dropout(m); m.train()
opt = bbbp_b200.AdamW(m.parameters(), lr=1e-4, weight_decay=1e-5); crit = bbbp_b200.BCEWithLogitsLoss()
fp, img, y = torch.randn(32, 2048, device=dev), torch.randn(32, 49152, device=dev), (torch.rand(32, device=dev) < 0.64).float()
def eager():
    opt.zero_grad(); loss = crit(m(fp, img).squeeze(), y); loss.backward(); opt.step()
step = bbbp_b200.GraphedTrainStep(m, opt, crit)
print(f"Morgan-2048 train step batch 32 (fp32, BCE): eager {t(eager):.2f} ms, graph {t(lambda: step(fp, img, y)):.2f} ms", flush=True)
