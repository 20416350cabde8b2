import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, bbbp_b200
from oracle import nets
from conftest import seeded_inputs
IMG = 49152
def run(mode, fork=True, prio=True, wfork=True, imfork=True, batches=(32, 32, 32)):
    torch.manual_seed(4)
    ref = nets.build("tcnn", 167, 128); model = bbbp_b200.build("tcnn", 167, 128); model.load_state_dict(ref.state_dict()); model.cuda()
    nets.zero_dropout(model); model.train()
    opt = bbbp_b200.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5); crit = bbbp_b200.MSELoss()
    step = bbbp_b200.GraphedTrainStep(model, opt, crit, fork_image_branch=fork); step.high_priority_chain = prio; step.fork_weight_grads = wfork; step.fork_image = imfork
    out = []
    for i, b in enumerate(batches):
        fp, img, y = (t.cuda() for t in seeded_inputs(900 + i, b, 167, IMG))
        if mode == "graph": loss = step(fp, img, y)
        else:
            opt.zero_grad(); loss = crit(model(fp, img).squeeze(), y); loss.backward(); opt.step()
        g = torch.cat([p.grad.detach().reshape(-1) for p in model.parameters()]).double()
        w = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).double()
        out.append((float(loss.detach()), float(g.abs().sum()), float(w.abs().sum())))
    gr = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    return out, gr
res = {}
for name, args in (("eager", ("eager",)), ("prio both", ("graph", True, True, True, True)), ("prio both 2", ("graph", True, True, True, True))):
    res[name] = run(*args, batches=(32,32,10,32))
    print(name, res[name][0])
for name in res:
    if name == "eager": continue
    bad = [(k, float((v - res[name][1][k]).abs().max() / (v.abs().max() + 1e-30))) for k, v in res["eager"][1].items() if not torch.equal(v, res[name][1][k])]
    print(name, "DIFF:", [(k.replace("fingerprint_transformer.layers.", "L"), "%.1e" % e) for k, e in bad])
