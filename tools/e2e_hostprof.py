"""Where does the host time of predict_from_host go?  cProfile of 8 calls at a small chunk size + device timeline gaps."""
import os, sys, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
dev = torch.device("cuda:0"); torch.manual_seed(0)
m = bbbp_b200.MixedInputModel(167, 128).to(dev).eval().set_precision("bf16")
n = 8192; chunk = int(os.environ.get("CHUNK", 512))
packed = torch.randint(0, 256, (n, 21), dtype=torch.uint8).pin_memory()
img8 = torch.randint(0, 256, (n, 3, 128, 128), dtype=torch.uint8).pin_memory()
out_host = torch.empty(n, dtype=torch.float32).pin_memory()
fn = lambda: m.predict_from_host(packed, img8, 256, chunk_molecules=chunk, packed=True, out_host=out_host)
for _ in range(3): fn()
torch.cuda.synchronize()
t0 = time.perf_counter(); fn(); t_host = time.perf_counter() - t0; torch.cuda.synchronize(); t_all = time.perf_counter() - t0
print(f"chunk {chunk}: host issue time {t_host*1e3:.2f} ms, until device idle {t_all*1e3:.2f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(4): fn()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
