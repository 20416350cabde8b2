"""Device timeline of one resident screening step (8192 molecules, bf16 mode): per-kernel time in launch order."""
import os, sys, json, tempfile, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0"); torch.manual_seed(0)
F = int(os.environ.get("F", 167)); variant = os.environ.get("VARIANT", "tcnn")
m = bbbp_b200.build(variant, F, 128).to(dev).eval().set_precision(os.environ.get("PREC", "bf16"))
n = int(os.environ.get("N", 8192))
fp, img = torch.randn(n, F, device=dev), torch.randn(n, 49152, device=dev)
with torch.no_grad():
    for _ in range(3): m.predict_batches(fp, img, 256, max_rows_per_pass=n)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m.predict_batches(fp, img, 256, max_rows_per_pass=n)
        torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace_inf.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in ev)
print(f"span {t1 - t0:.1f} us, {len(ev)} ops, busy {sum(e['dur'] for e in ev):.1f}")
lines = [f"{e['ts'] - t0:9.1f} {e['dur']:8.1f} {e['name'][:90]}" for e in ev]
os.makedirs("gpurun_out", exist_ok=True)
open(f"gpurun_out/infer_timeline_{variant}_{F}.txt", "w").write("\n".join(lines))
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    k = e["name"].split("(")[0][-60:]; agg[k][0] += 1; agg[k][1] += e["dur"]
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:14]: print(f"{v:9.1f} {c:4d} {v / c:8.1f} {k}")
