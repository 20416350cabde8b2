"""Device timeline of ONE CUDA-graph replay of the train step (torch.profiler / CUPTI): per-stream busy time and the
longest kernels, to see which branch of the captured graph is the critical path."""
import os, sys, json, tempfile, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200

from torch.profiler import profile, ProfilerActivity

B = int(os.environ.get("B", 32)); prec = os.environ.get("PREC", "fp32")
dev = torch.device("cuda:0"); torch.manual_seed(0)
m = bbbp_b200.MixedInputModel(167, 128).to(dev); bbbp_b200.zero_dropout(m); m.train().set_precision(prec)
opt = bbbp_b200.AdamW(m.parameters(), lr=1e-4, weight_decay=1e-5); crit = bbbp_b200.MSELoss()
fp, img, y = torch.randn(B, 167, device=dev), torch.randn(B, 49152, device=dev), torch.randn(B, device=dev)
step = bbbp_b200.GraphedTrainStep(m, opt, crit, fork_image_branch=os.environ.get("FORK", "1") == "1")
for _ in range(5): step(fp, img, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step(fp, img, y)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
# split into replays at the adamw kernel
ends = [i for i, e in enumerate(ev) if "adamw" in e["name"]]
a, b = ends[-2] + 1, ends[-1] + 1
one = ev[a:b]
t0 = one[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in one)
print(f"replay span {t1 - t0:.1f} us, {len(one)} device ops")
by_stream = collections.defaultdict(list)
for e in one: by_stream[e["args"].get("stream")].append(e)
for s, L in by_stream.items():
    busy = sum(e["dur"] for e in L)
    print(f"stream {s}: {len(L)} ops, busy {busy:.1f} us, from {L[0]['ts'] - t0:.1f} to {max(e['ts'] + e['dur'] for e in L) - t0:.1f}")
out = []
for e in one:
    out.append(f"{e['ts'] - t0:9.1f} {e['dur']:8.1f} s{e['args'].get('stream')} {e['name'][:70]}")
os.makedirs("gpurun_out", exist_ok=True)
open(f"gpurun_out/train_timeline_{prec}_b{B}.txt", "w").write("\n".join(out))
agg = collections.defaultdict(lambda: [0, 0.0])
for e in one:
    k = e["name"].split("(")[0][-50:]; agg[k][0] += 1; agg[k][1] += e["dur"]
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:25]: print(f"{v:8.1f} {c:4d} {v / c:7.1f} {k}")
