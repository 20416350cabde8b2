"""TMEM -> register read bandwidth of one SM (tcgen05.ld), the floor of the pooled-convolution epilogues."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bbbp_b200
from bbbp_b200._lib import lib, check
out = torch.zeros(3, dtype=torch.int64, device="cuda")
for warps in (4, 8):
    for reps in (200, 2000):
        check(lib.bbbp_debug_tmem_read_probe(out.data_ptr(), warps, reps, torch.cuda.current_stream().cuda_stream), "probe")
        torch.cuda.synchronize()
        cyc, byt, _ = out.tolist()
        print(f"tcgen05.ld 32x32b.x32, {warps} warps, {reps} reps: {byt} B in {cyc} cycles = {byt / cyc:.1f} B/clk/SM "
              f"-> {byt / cyc * 148 * 1.965e9 / 1e12:.1f} TB/s chip-wide at 1965 MHz")
# conv1: 4 window members x 32 channels x 16384 pre-pool/4 pooled pixels x 4 B = 2 MiB of accumulators per molecule
