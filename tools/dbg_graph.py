import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, bbbp_b200
from oracle import nets
from conftest import seeded_inputs
IMG = 49152
def run(mode, fork=True, batches=(32, 32, 32, 10, 32), lr_change=True):
    torch.manual_seed(4)
    ref = nets.build("tcnn", 167, 128); model = bbbp_b200.build("tcnn", 167, 128); model.load_state_dict(ref.state_dict()); model.cuda()
    nets.zero_dropout(model); model.train()
    opt = bbbp_b200.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5); crit = bbbp_b200.MSELoss()
    step = bbbp_b200.GraphedTrainStep(model, opt, crit, fork_image_branch=fork)
    out = []
    for i, b in enumerate(batches):
        if i == 2 and lr_change: opt.param_groups[0]["lr"] = 3e-5
        fp, img, y = (t.cuda() for t in seeded_inputs(900 + i, b, 167, IMG))
        if mode == "graph": loss = step(fp, img, y)
        else:
            opt.zero_grad(); loss = crit(model(fp, img).squeeze(), y); loss.backward(); opt.step()
        w = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).double()
        out.append((float(loss.detach()), float(w.sum()), float(w.abs().sum())))
    return out
a = run("eager"); b = run("eager"); c = run("graph", True); d = run("graph", False)
e = run("eager", lr_change=False); f = run("graph", True, lr_change=False)
for name, r in (("eager", a), ("eager2", b), ("graph fork", c), ("graph nofork", d), ("eager nolr", e), ("graph nolr", f)):
    print(name)
    for x in r: print("   %.10f %.12f %.10f" % x)
g = run("eager", batches=(10, 10)); h = run("graph", batches=(10, 10))
print("b10 first", g, h)
