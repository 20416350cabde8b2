#!/usr/bin/env python
"""Throughput bench of the hot path: batched inference of the multi-input transformer-CNN
(MixedInputModel, MACCS 167-bit + 3x128x128 depiction) with the reference's batch-256 semantics.

    python bench.py --gpus N --steps K --warmup W            # our arm (N>1 under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU PyTorch path (oracle port)

One "step" scores ``--groups`` independent reference batches of 256 molecules (default 64 -> 16 384
molecules; inputs 3.2 GB fp32 per step, far larger than the 126 MB L2, so nothing is cache-resident
between iterations).  ``value`` times the step with inputs already in HBM; ``e2e`` times the same call
from pinned HOST buffers including the H2D copy of the step's inputs and the D2H read of its scores.
Prints ONE JSON line on rank 0.

Precision: the headline (``value``, ``e2e``, ``roofline``) is the STRICT tensor-core mode -- fp16 operands with hi + lo
activation pairs, |d logBB| <= 1e-3 against the fp32 reference at trained output scale
(tests/test_trained_parity_gpu.py) -- i.e. the reference's own arithmetic class.  The faster one-pass modes are reported
beside it under ``by_precision`` with their own stated tolerances.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_BITS, IMG = 167, 3 * 128 * 128
BATCH = 256
METRIC, UNIT = "molecules/sec multi-input NN inference", "molecules/s"
# algorithmic work (SURVEY 8d / DESIGN.md): conv2 = 75.50 MMAC per molecule
CONV2_FLOP_PER_MOL = 2 * 64 * 64 * 64 * 288
# DRAM bytes of the conv2 kernel per molecule from the committed `ncu --set full` capture
# (profiles/r01_ncu_conv_umma_full.txt: dram__bytes_read 2.186054 GB + dram__bytes_write 1.046418 GB per 8192-molecule
# launch); the algorithmic figure is 256 KiB in + 128 KiB out = 393 216 B, i.e. no re-reads
CONV2_DRAM_BYTES_PER_MOL = (2.186054e9 + 1.046418e9) / 8192
# ... and of its strict-mode instantiation (profiles/r02_ncu_conv.txt: 2.148330 GB + 1.042366 GB per 4 096-molecule launch);
# algorithmic: (hi, lo) pairs in and out = 2 x (256 KiB + 128 KiB) = 786 432 B, i.e. no re-reads either
CONV2_STRICT_DRAM_BYTES_PER_MOL = (2.148330e9 + 1.042366e9) / 4096
# ... and of the background-referenced strict instantiation, one pass over single fp16 tensors (profiles/r02_ncu_conv_ffn_flash.txt:
# 1.076648 GB + 0.506714 GB per 4 096-molecule launch = 386.6 KB per molecule against 393 216 B algorithmic: no re-reads)
CONV2_BG_DRAM_BYTES_PER_MOL = (1.076648e9 + 0.506714e9) / 4096
# first layer (conv1 + ReLU + pool): HBM-bound by its arithmetic intensity (28.3 MFLOP over 458 752 B = 62 flop/B against a ridge
# of ~208).  Algorithmic bytes per molecule: the planar input (fp32 196 608 B, or uint8 49 152 B) + the (64, 64, 32) 16-bit output
CONV1_OUT_BYTES_PER_MOL = 64 * 64 * 32 * 2
# DRAM bytes per molecule of the uint8 / exact-integer instantiation (profiles/r02_ncu_conv_ffn_flash.txt: 0.203347 GB read +
# 1.022220 GB written per 4 096-molecule launch = 299 KB against 311 296 B algorithmic)
CONV1_U8_DRAM_BYTES_PER_MOL = (0.203347e9 + 1.022220e9) / 4096
# ... and of the fp32-plane strict instantiation, (hi, lo) staging (profiles/r02_ncu_conv_fp32_planes.txt: 0.812390 GB + 1.032090 GB
# per 4 096-molecule launch = 450 KB against 458 752 B algorithmic: no re-reads)
CONV1_F32_DRAM_BYTES_PER_MOL = (0.812390e9 + 1.032090e9) / 4096
FWD_FLOP_PER_MOL = 207.2e6
IN_BYTES_PER_MOL = (F_BITS + IMG + 1) * 4
# the workload both arms run (identical dict in both JSON lines; per-arm sample sizes live outside it)
CONFIG = {"workload": "screen_maccs_b256", "model": "MixedInputModel(167,128) 20250113 (13.46M params)", "batch": BATCH,
          "inputs": "MACCS-167 fingerprint + 3x128x128 depiction per molecule"}
# molecules per staged chunk of the host pipeline: a chunk's replay costs ~0.55 ms + its compute, its copy 0.94 ms per 1 024
# (tools/e2e_sweep.py on B200, strict mode, 16 384 molecules: 1 024 -> 20.0 ms, 2 048 -> 16.4 ms, 4 096 -> 17.5 ms)
E2E_CHUNK = 2048
# sparse depictions cut the copy to ~5 KB per molecule: the link outruns the arithmetic 5:1, so only the FIRST chunk's copy is
# exposed and every later copy hides behind the chunk before it -- a short first chunk, then one long one (a schedule of chunk
# lengths).  SPARSE=1 SCHEDULES=1 tools/e2e_sweep.py, strict, 16 384 molecules: equal chunks of 8 192 -> 7.68 ms, one chunk of
# 16 384 -> 7.98, (2 048, 14 336) -> 7.27, (4 096, 12 288) -> 7.38, (2 048, 6 144, 8 192) -> 7.74 (profiles/r02_e2e_sweep_sparse_strict.txt)
E2E_CHUNK_SPARSE = (2048, 14336)
# (four ranks behind one host bridge, tools/gpu_r2_multi3.sh: (2 048, 14 336) 7.19 ms, (4 096, 12 288) 7.12, (2 048, 6 144, 8 192) 7.35
# against 7.10 resident -- the schedule holds when ranks share a link, profiles/r02_e2e_schedules_n4.txt)
DTYPE = {"fp32": "f32", "bf16": "bf16", "fp16": "f16", "strict": "f16 (background-referenced activations, split small GEMMs, fp32 accumulate)"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return {"hbm": p["hbm_gbs"], "tensor_burst": p["bf16_tflops"], "tensor": p["bf16_tflops_sustained"], "src": "measured"}
    except Exception:
        return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "src": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (every ~5 ms; the
    timed region of a default run is only tens of milliseconds, too short for `nvidia-smi -lms 200`)."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index=0):
        self.index, self.samples, self.reasons, self.thread, self.stop, self.err = index, [], 0, None, False, None
        self.max_mhz, self.power, self.nvml, self.h = None, [], None, None
        try:                       # NVML start-up costs ~100 ms: do it here, before any timed region
            import pynvml
            import torch
            pynvml.nvmlInit()
            bus = torch.cuda.get_device_properties(index).pci_bus_id   # CUDA_VISIBLE_DEVICES may renumber devices
            for i in range(pynvml.nvmlDeviceGetCount()):
                hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                if int(pynvml.nvmlDeviceGetPciInfo(hi).bus) == int(bus):
                    self.h = hi
            self.h = self.h or pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception as e:     # no NVML (CPU container)
            self.err = str(e)[:100]

    def _sample(self):
        n = self.nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        self.reasons |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        self.power.append(n.nvmlDeviceGetPowerUsage(self.h) / 1000.0)

    def _run(self):
        try:
            while not self.stop:
                self._sample()
                time.sleep(0.003)
        except Exception as e:
            self.err = str(e)[:100]

    def __enter__(self):
        import threading
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.thread is not None:
            self.thread.join(timeout=2)
            try:
                self._sample()         # one more sample while the last step's kernels are still draining
            except Exception:
                pass

    def summary(self):
        out = {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": [n for n, bit in self.BAD.items() if self.reasons & bit], "samples": len(self.samples),
               "power_w_max": max(self.power) if self.power else None}
        if self.err:
            out["error"] = self.err
        return out


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML reports as local to GPU ``index`` BEFORE any pinned host buffer is allocated, so
    the staging memory is first-touched on the GPU's own NUMA node: with 8 ranks streaming 55 GB/s each, host buffers on
    the far socket halve the H2D rate.  Returns the number of CPUs in the mask (None when NVML / affinity is unavailable)."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        handle = None
        for i in range(pynvml.nvmlDeviceGetCount()):
            hi = pynvml.nvmlDeviceGetHandleByIndex(i)
            if int(pynvml.nvmlDeviceGetPciInfo(hi).bus) == int(bus):
                handle = hi
        if handle is None:
            return None
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def synthetic_inputs(n, seed, device):
    """B3DB/ZINC-shaped synthetic molecules: MACCS bits ~ Bernoulli(0.25) with bit 0 = 0, per-molecule
    z-scored (P1); depictions ~ N(0,1) (a z-scored image, P2).  Generated with torch RNG, outside any timer."""
    import torch
    g = torch.Generator(device=device).manual_seed(20250113 + seed)
    bits = (torch.rand(n, F_BITS, generator=g, device=device) < 0.25).float()
    bits[:, 0] = 0
    mean = bits.mean(1, keepdim=True)
    std = bits.std(1, unbiased=False, keepdim=True).clamp_min(1e-12)
    fp = ((bits - mean) / std).contiguous()
    img = torch.randn(n, IMG, generator=g, device=device)
    return fp, img


def cpu_reference_rate(n_batches, threads, warm=1):
    """The reference's CPU PyTorch path (oracle/nets.py, pinned to the reference classes): eval, no_grad,
    fp32, inputs resident as contiguous CPU tensors, batch 256.  Returns (mol/s, seconds, molecules)."""
    import torch
    from oracle import nets
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = nets.build("tcnn", F_BITS, 128).eval()
    fp, img = synthetic_inputs(BATCH * max(1, n_batches), 7, "cpu")
    with torch.no_grad():
        for _ in range(warm):
            model(fp[:BATCH], img[:BATCH])
        t0 = time.perf_counter()
        for b in range(n_batches):
            model(fp[b * BATCH:(b + 1) * BATCH], img[b * BATCH:(b + 1) * BATCH])
        dt = time.perf_counter() - t0
    return n_batches * BATCH / dt, dt, n_batches * BATCH


def train_step_ms(torch, bbbp_b200, nets, dev, with_cpu=True):
    """Secondary figure of BASELINE.json's metric ("train step ms"): fwd + bwd + AdamW of the same network with the
    reference's settings (20250113.py:172,187-191: batch 32, AdamW lr 1e-4 wd 1e-5, MSE), fp32 kernels, dropout off
    (the reference's dominant regime).  CUDA events, 3 warm-ups, inputs resident.  ``ms_batch*`` is the
    public training API (bbbp_b200.GraphedTrainStep: the whole step as one CUDA-graph replay, bit-identical to the
    eager loop body); ``eager_ms_batch*`` is the reference's loop body issued launch by launch from Python."""
    out = {"precision": "fp32", "optimizer": "bbbp_b200.AdamW (one fused launch)", "loss": "bbbp_b200.MSELoss",
           "api": "bbbp_b200.GraphedTrainStep(model, optimizer, criterion)(fingerprint, image, target)"}
    torch.manual_seed(0)
    model = bbbp_b200.MixedInputModel(F_BITS, 128).to(dev)
    bbbp_b200.zero_dropout(model)
    model.train()
    opt = bbbp_b200.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)
    crit = bbbp_b200.MSELoss()
    graphed = bbbp_b200.GraphedTrainStep(model, opt, crit)

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for batch in (32, 256):
        fp, img = synthetic_inputs(batch, 3, dev)
        y = torch.randn(batch, device=dev) * 0.75 - 0.1

        def eager():
            opt.zero_grad()
            loss = crit(model(fp, img).squeeze(), y)
            loss.backward()
            opt.step()
        out[f"eager_ms_batch{batch}"] = timed(eager, 10)
        out[f"ms_batch{batch}"] = timed(lambda: graphed(fp, img, y), 20)
    # tensor-core training mode (bf16 operands): every nn.Linear forward and, at batch >= 64, forward AND backward of the
    # image branch on tcgen05 (autograd.ImageBranchTensorCore); the reference's batch 32 stays latency-bound either way
    model.set_precision("bf16")
    fp, img = synthetic_inputs(256, 3, dev)
    y = torch.randn(256, device=dev) * 0.75 - 0.1
    out["ms_batch256_tensor_core_bf16"] = timed(lambda: graphed(fp, img, y), 20)
    model.set_precision("fp32")
    if with_cpu:
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        torch.manual_seed(0)
        ref = nets.zero_dropout(nets.build("tcnn", F_BITS, 128)).train()
        ropt = torch.optim.AdamW(ref.parameters(), lr=1e-4, weight_decay=1e-5)
        fp, img = synthetic_inputs(32, 3, "cpu")
        y = torch.randn(32) * 0.75 - 0.1
        nets.train_step(ref, ropt, fp, img, y)
        t0 = time.perf_counter()
        for _ in range(3):
            nets.train_step(ref, ropt, fp, img, y)
        out["cpu_ms_batch32"] = (time.perf_counter() - t0) / 3 * 1e3
        out["cpu_cores"] = threads
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    per_step = 2                                   # bounded sample: 2 reference batches of 256 per step
    for _ in range(max(0, args.warmup)):
        cpu_reference_rate(1, threads, warm=0)
    t_all, n_all = 0.0, 0
    for _ in range(args.steps):
        _, dt, n = cpu_reference_rate(per_step, threads, warm=0)
        t_all += dt
        n_all += n
    v = n_all / t_all
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": CONFIG, "molecules_per_step": per_step * BATCH,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{args.steps} steps x {per_step} batches of {BATCH} molecules, torch {torch.__version__} CPU fp32, "
                                       "oracle/nets.py (pinned to the reference classes)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--groups", type=int, default=64, help="reference batches of 256 molecules per step (64 -> 16 384 molecules)")
    ap.add_argument("--precision", default=os.environ.get("BBBP_BENCH_PRECISION", "strict"),
                    choices=["fp32", "bf16", "fp16", "strict"], help="precision mode of the headline numbers")
    ap.add_argument("--also", default="fp16,bf16", help="comma-separated extra modes timed resident + e2e (by_precision)")
    ap.add_argument("--cpu-batches", type=int, default=48, help="bounded CPU-baseline sample (batches of 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary train-step measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import bbbp_b200
    from bbbp_b200 import ops
    from oracle import nets          # cpu_baseline leg and weight init only (never on the timed GPU path)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.require_device()
    affinity = bind_to_gpu_numa_node(local) if world > 1 else None

    torch.manual_seed(0)
    model = bbbp_b200.MixedInputModel(F_BITS, 128)       # random-init weights of the reference architecture
    model.to(dev).eval()
    model.set_precision(args.precision)
    n = args.groups * BATCH
    fp, img = synthetic_inputs(n, rank, dev)
    fp_host = fp.cpu().pin_memory()
    img_host = img.cpu().pin_memory()
    scores_host = torch.empty(n, dtype=torch.float32).pin_memory()
    # the same kind of molecules in the compact screening formats (SURVEY cfg4): packed MACCS bits + uint8 depictions
    g = torch.Generator().manual_seed(99 + rank)
    bits = (torch.rand(n, F_BITS, generator=g) < 0.25)
    bits[:, 0] = False
    weights = (2 ** torch.arange(8)).to(torch.uint8)
    padded = torch.zeros(n, (F_BITS + 7) // 8 * 8, dtype=torch.bool)
    padded[:, :F_BITS] = bits
    packed_host = (padded.view(n, -1, 8).to(torch.uint8) * weights).sum(-1).to(torch.uint8).pin_memory()
    img8_host = torch.full((n, 3, 128, 128), 255, dtype=torch.uint8)
    strokes = torch.rand(n, 1, 128, 128, generator=g) < 0.06
    img8_host[strokes.expand(-1, 3, -1, -1)] = 40
    img8_host = img8_host.pin_memory()
    # the same depictions in the lossless sparse encoding (bit mask of the non-white pixels + their RGB triples)
    sparse_host = bbbp_b200.SparseDepictions.encode(img8_host.numpy())
    gathered = torch.empty(world * n, device=dev, dtype=torch.float32) if world > 1 else None

    def step_resident():
        s = model.predict_batches(fp, img, BATCH, max_rows_per_pass=n)
        if world > 1:
            dist.all_gather_into_tensor(gathered, s)        # the one collective: score gather over NVLink
        return s

    def step_e2e():
        f = fp_host.to(dev, non_blocking=True)
        i = img_host.to(dev, non_blocking=True)
        s = model.predict_batches(f, i, BATCH, max_rows_per_pass=n)
        if world > 1:
            dist.all_gather_into_tensor(gathered, s)
        scores_host.copy_(s, non_blocking=True)
        return s

    def step_e2e_compact():
        # the user-facing host API: chunked H2D on a copy stream overlapped with scoring, D2H of the scores, and a wait
        # for that copy -- every step ends with its scores readable in host memory
        _, s = model.predict_from_host(packed_host, img8_host, BATCH, chunk_molecules=E2E_CHUNK, packed=True, out_host=scores_host,
                                       return_device=True)
        if world > 1:
            dist.all_gather_into_tensor(gathered, s)

    sparse_schedule = E2E_CHUNK_SPARSE
    if os.environ.get("BBBP_E2E_SCHEDULE"):          # measurement override, e.g. "2048,14336"
        sparse_schedule = tuple(int(c) for c in os.environ["BBBP_E2E_SCHEDULE"].split(","))

    def step_e2e_sparse():
        _, s = model.predict_from_host(packed_host, sparse_host, BATCH, chunk_molecules=sparse_schedule, packed=True, out_host=scores_host,
                                       return_device=True)
        if world > 1:
            dist.all_gather_into_tensor(gathered, s)

    scores_host2 = torch.empty(n, dtype=torch.float32).pin_memory()
    stream_state = [0]

    def step_e2e_streaming():
        # the same API in its streaming mode (synchronize=False, alternating result buffers): consecutive shards overlap,
        # the first H2D chunks of step k+1 fly while the last chunk of step k is scored; results are complete at the
        # synchronisation that closes the timed region
        stream_state[0] ^= 1
        _, s = model.predict_from_host(packed_host, img8_host, BATCH, chunk_molecules=E2E_CHUNK, packed=True,
                                       out_host=scores_host2 if stream_state[0] else scores_host, return_device=True,
                                       synchronize=False)
        if world > 1:
            dist.all_gather_into_tensor(gathered, s)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    with torch.no_grad():
        for _ in range(args.warmup):
            step_resident()
        ops.KERNEL_TIMER.enable({"conv2", "conv1"})
        l0 = bbbp_b200._lib.lib.bbbp_launch_count()
        with ClockSampler(local) as clocks:
            ms = timed(step_resident, args.steps)
        launches = bbbp_b200._lib.lib.bbbp_launch_count() - l0
        conv2_ms, conv2_mols = ops.KERNEL_TIMER.collect("conv2")
        conv1_ms, _ = ops.KERNEL_TIMER.collect("conv1")
        ops.KERNEL_TIMER.disable()
        for _ in range(2):
            step_e2e()
        ms_e2e32 = timed(step_e2e, args.steps)
        for _ in range(2):
            step_e2e_compact()
        ms_e2e = timed(step_e2e_compact, args.steps)
        for _ in range(2):
            step_e2e_streaming()
        ms_e2e_stream = timed(step_e2e_streaming, args.steps)
        for _ in range(2):
            step_e2e_sparse()
        ms_e2e_sparse = timed(step_e2e_sparse, args.steps)
        by_precision = {}
        for mode in [m for m in args.also.split(",") if m and m != args.precision]:
            model.set_precision(mode)
            for _ in range(3):
                step_resident()
            m_res = timed(step_resident, args.steps)
            for _ in range(2):
                step_e2e_sparse()          # (captures this mode's chunk graphs outside the timed region)
            m_e2e = timed(step_e2e_sparse, args.steps)
            by_precision[mode] = {"value": world * n * args.steps / (m_res * 1e-3), "e2e": world * n * args.steps / (m_e2e * 1e-3),
                                  "unit": UNIT, "dtype": DTYPE[mode], "ms_per_step": m_res / args.steps}
        model.set_precision(args.precision)

    total_mols = world * n * args.steps
    value = total_mols / (ms * 1e-3)
    e2e_dense = total_mols / (ms_e2e * 1e-3)
    e2e = total_mols / (ms_e2e_sparse * 1e-3)
    e2e32 = total_mols / (ms_e2e32 * 1e-3)
    sparse_bytes = sparse_host.nbytes() + packed_host.numel()
    pk = peaks()
    roof = None
    if conv2_ms:
        per_launch_ms = statistics.mean(conv2_ms)
        flop = CONV2_FLOP_PER_MOL * statistics.mean(conv2_mols)
        ach = flop / (per_launch_ms * 1e-3) / 1e12
        # strict mode: ONE pass over background-referenced fp16 activations (model.strict_background, the default); its
        # first form carried (hi, lo) pairs through two MMAs per K step and twice the bytes
        passes = 2 if (args.precision == "strict" and not model.strict_background) else 1
        bg = args.precision == "strict" and model.strict_background
        traffic = (CONV2_STRICT_DRAM_BYTES_PER_MOL if passes == 2 else CONV2_BG_DRAM_BYTES_PER_MOL if bg else
                   CONV2_DRAM_BYTES_PER_MOL) * statistics.mean(conv2_mols)
        roof = {"kernel": "conv2 (3x3, 32->64, +bias+ReLU+maxpool) implicit GEMM", "bound": "tensor", "achieved": ach,
                "peak": pk["tensor"], "unit": "TFLOP/s", "frac": ach / pk["tensor"],
                "traffic": traffic, "traffic_unit": "bytes/launch",
                "traffic_source": ("ncu --set full of the pair-mode kernel (profiles/r02_ncu_conv.txt: tensor pipe 59 %, DRAM = the "
                                   "algorithmic bytes of the (hi, lo) pairs)" if passes == 2 else
                                   "ncu --set full of the background-referenced strict kernel (profiles/r02_ncu_conv_ffn_flash.txt: tensor "
                                   "pipe 63 %, DRAM 1.58 GB per 4 096 molecules = the algorithmic bytes)" if bg else
                                   "ncu --set full of the one-pass kernel (profiles/r01_ncu_conv_umma_full.txt)") + ", scaled to this launch's molecules",
                "peak_source": pk["src"] + " bf16_tflops_sustained", "launch_ms": per_launch_ms,
                "mma_passes": passes, "executed_tflops": ach * passes, "executed_frac": ach * passes / pk["tensor"],
                "note": "achieved = ALGORITHMIC flops (151.0 MFLOP per molecule, SURVEY 8d) / CUDA-event time; executed_* multiplies by "
                        "the MMA passes the mode issues (1, or 2 in the pair form of the strict mode)",
                "share_of_step": sum(conv2_ms) / ms, "whole_model_tflops": FWD_FLOP_PER_MOL * value / 1e12,
                "conv1_launch_ms": statistics.mean(conv1_ms) if conv1_ms else None}
    # the first layer, against the HBM roofline that bounds it.  Whichever of the two convolution kernels took more of the timed
    # region is reported under "roofline" (strict mode from fp32 planes: the first layer, whose (hi, lo) staging makes it the
    # longest kernel of the step; one-pass modes: conv2); the other one under "roofline_second"
    roof1 = None
    if conv1_ms and conv2_ms:
        t1 = statistics.mean(conv1_ms)
        bytes1 = (IMG * 4 + CONV1_OUT_BYTES_PER_MOL) * statistics.mean(conv2_mols)
        gbs = bytes1 / (t1 * 1e-3) / 1e9
        roof1 = {"kernel": "conv1 (3x3, 3->32, +bias+ReLU+maxpool) implicit GEMM straight from the planar fp32 input", "bound": "hbm",
                 "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                 "traffic": CONV1_F32_DRAM_BYTES_PER_MOL * statistics.mean(conv2_mols), "traffic_unit": "bytes/launch",
                 "traffic_source": "ncu --set full of the strict-mode launch from fp32 planes (profiles/r02_ncu_conv_fp32_planes.txt: 450 KB "
                                   "per molecule = the algorithmic bytes; one-pass modes move the same tensors), scaled to this launch's "
                                   f"molecules; the uint8 / exact-integer instantiation moves {CONV1_U8_DRAM_BYTES_PER_MOL:.0f} B per molecule "
                                   "against 311 296 B algorithmic (profiles/r02_ncu_conv_ffn_flash.txt)",
                 "peak_source": pk["src"] + " hbm_gbs", "launch_ms": t1, "share_of_step": sum(conv1_ms) / ms,
                 "note": "achieved = ALGORITHMIC bytes (196 608 B fp32 planes in + 262 144 B 16-bit NHWC out per molecule) / CUDA-event "
                         "time; the kernel is bound by its producers / epilogue warps, not by DRAM (DESIGN.md section 5b item 1)"}
        if sum(conv1_ms) > sum(conv2_ms):
            roof, roof1 = roof1, roof
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE[args.precision], "data": "synthetic",
            "config": CONFIG, "molecules_per_step": world * n, "molecules_per_step_per_gpu": n,
            "resident_input": "fp32 z-scored fingerprint + fp32 CHW image (the reference's input contract)",
            "l2_policy": f"inputs {n * IN_BYTES_PER_MOL / 1e6:.0f} MB per step > 126 MB L2", "precision": args.precision,
            "parity": "strict: |d logBB| <= 1e-3 vs the fp32 reference at trained output scale; fp16 <= 1e-2 x spread; bf16 <= 9e-2 x "
                      "spread (tests/test_trained_parity_gpu.py)",
            "by_precision": by_precision,
            "clocks": clocks.summary(), "gpu_launches": int(launches),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(sparse_bytes), "d2h_bytes_per_step": n * 4,
                    "ms_per_step": ms_e2e_sparse / args.steps, "chunk_schedule": list(sparse_schedule),
                    "api": "model.predict_from_host(packed MACCS bits uint8, bbbp_b200.SparseDepictions, packed=True): pinned host -> "
                           "chunked H2D overlapped with [sparse decode to uint8 CHW + unpack + z-score + in-kernel image "
                           "normalisation + forward] -> D2H scores, readable in host memory when the call returns",
                    "input_format": "lossless sparse depictions: 2 048-byte bit mask of the non-white pixels + their RGB triples "
                                    f"({sparse_bytes / n:.0f} B per molecule here, 6 084 B on the real B3DB depictions; dense uint8 = 49 152 B)"},
            "e2e_dense_u8": {"value": e2e_dense, "unit": UNIT, "h2d_bytes_per_step": n * ((F_BITS + 7) // 8 + IMG), "d2h_bytes_per_step": n * 4,
                             "ms_per_step": ms_e2e / args.steps,
                             "api": "the same call with dense uint8 CHW depictions (round 1's e2e format): PCIe-bound at 55 GB/s"},
            "e2e_streaming": {"value": total_mols / (ms_e2e_stream * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n * ((F_BITS + 7) // 8 + IMG),
                              "d2h_bytes_per_step": n * 4, "ms_per_step": ms_e2e_stream / args.steps,
                              "api": "the same call with synchronize=False on consecutive shards (per-slot events carry across calls; "
                                     "results read after the closing synchronisation)"},
            "e2e_fp32_contract": {"value": e2e32, "unit": UNIT, "h2d_bytes_per_step": n * (F_BITS + IMG) * 4,
                                  "d2h_bytes_per_step": n * 4, "ms_per_step": ms_e2e32 / args.steps,
                                  "api": "model.predict_batches(fp32 fingerprint, fp32 image) from pinned host buffers"},
            "roofline": roof, "roofline_second": roof1}
    if affinity is not None:
        line["host_cpus_local_to_gpu"] = affinity
    if world == 1 and not args.no_train:
        line["train_step"] = train_step_ms(torch, bbbp_b200, nets, dev, with_cpu=not args.no_cpu_baseline)
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            v, dt, mols = cpu_reference_rate(args.cpu_batches, threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{mols} molecules ({args.cpu_batches} batches of {BATCH}) in {dt:.1f} s, torch "
                                              f"{torch.__version__} CPU fp32, oracle/nets.py restatement of the reference classes"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
