"""Train-step timing: the reference loop body eagerly vs one CUDA-graph replay (GraphedTrainStep)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200


dev = torch.device("cuda:0")
out = {}
for prec in os.environ.get("PRECS", "fp32,bf16").split(","):
    for B in [int(b) for b in os.environ.get("BATCHES", "32,256").split(",")]:
        torch.manual_seed(0)
        m = bbbp_b200.MixedInputModel(167, 128).to(dev); bbbp_b200.zero_dropout(m); m.train().set_precision(prec)
        opt = bbbp_b200.AdamW(m.parameters(), lr=1e-4, weight_decay=1e-5); crit = bbbp_b200.MSELoss()
        fp, img, y = torch.randn(B, 167, device=dev), torch.randn(B, 49152, device=dev), torch.randn(B, device=dev)

        def eager():
            opt.zero_grad(); loss = crit(m(fp, img).squeeze(), y); loss.backward(); opt.step(); return loss
        step = bbbp_b200.GraphedTrainStep(m, opt, crit, fork_image_branch=os.environ.get("FORK", "1") == "1")

        def timeit(fn, n=20):
            for _ in range(3): fn()
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n): fn()
            e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
        out[f"{prec}_b{B}_eager_ms"] = timeit(eager)
        out[f"{prec}_b{B}_graph_ms"] = timeit(lambda: step(fp, img, y))
        print(prec, B, out[f"{prec}_b{B}_eager_ms"], out[f"{prec}_b{B}_graph_ms"], flush=True)
print(json.dumps(out))
