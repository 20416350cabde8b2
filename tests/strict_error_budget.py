"""Where the strict mode's remaining error comes from, measured ON THE DEVICE (test infrastructure, not a test).

    python tests/strict_error_budget.py

Trains the oracle network as tests/test_trained_parity_gpu.py does, then compares the product's strict-mode intermediates
with the oracle's fp32 intermediates on the 1 058 real depictions: conv1 output (hi + lo pair), conv2 output, the image
feature vector after Linear(65536, 128), and the final scores with single branches switched to the fp32 kernels.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import precision_study as ps  # noqa: E402


def main():
    import bbbp_b200
    from bbbp_b200 import ops, autograd as ag
    dev = torch.device("cuda:0")
    model, fp, img, y = ps.trained_state(200)
    sd = model.state_dict()
    ours = bbbp_b200.build("tcnn", 167, 128)
    ours.load_state_dict(sd, strict=True)
    ours.to(dev).eval()
    n = 256
    x = img[:n].view(n, 3, 128, 128)
    with torch.no_grad():
        r1 = F.max_pool2d(F.relu(F.conv2d(x.double(), sd["image_cnn.0.weight"].double(), sd["image_cnn.0.bias"].double(), padding=1)), 2)
        r2 = F.max_pool2d(F.relu(F.conv2d(r1, sd["image_cnn.3.weight"].double(), sd["image_cnn.3.bias"].double(), padding=1)), 2)
        r3 = F.relu(r2.flatten(1) @ sd["image_cnn.7.weight"].double().T + sd["image_cnn.7.bias"].double())
    conv1, conv2, fc = ours.image_cnn[0], ours.image_cnn[3], ours.image_cnn[7]
    rep = lambda name, got, ref: print(f"{name:58s} max {float((got.double().cpu() - ref).abs().max()):.3e}  "
                                       f"rel-to-max {float((got.double().cpu() - ref).abs().max() / ref.abs().max()):.2e}")
    with torch.no_grad():
        for mode in ("strict", "fp16"):
            fmt, split = ag.TENSOR_CORE[mode]
            w1 = ops.conv3x3_prepare_bf16(conv1.weight.detach(), fmt)
            w2 = ops.conv3x3_prepare_bf16(conv2.weight.detach(), fmt)
            wfc = ops.fc_weight_to_hwc_bf16(fc.weight.detach(), 64, 1024, fmt)
            xin = img[:n].to(dev)
            if split:
                y1, y1l = ops.conv1_from_image_bf16(xin, w1, conv1.bias, None, fmt=fmt, split=True)
                y2, y2l = ops.conv3x3_relu_pool_bf16(y1, w2, conv2.bias, 64, fmt=fmt, x_lo=y1l)
                o, _ = ops.gemm_bf16(y2.view(n, 65536), 65536, wfc, 128, bias=fc.bias, act="relu", split_k=8, fmt=fmt, a_lo=y2l.view(n, 65536))
                g1, g2 = y1.float() + y1l.float(), y2.float() + y2l.float()
            else:
                y1 = ops.conv1_from_image_bf16(xin, w1, conv1.bias, None, fmt=fmt)
                y2 = ops.conv3x3_relu_pool_bf16(y1, w2, conv2.bias, 64, fmt=fmt)
                o, _ = ops.gemm_bf16(y2.view(n, 65536), 65536, wfc, 128, bias=fc.bias, act="relu", split_k=8, fmt=fmt)
                g1, g2 = y1.float(), y2.float()
            rep(f"[{mode}] conv1 output (NHWC)", g1.permute(0, 3, 1, 2), r1)
            rep(f"[{mode}] conv2 output", g2.permute(0, 3, 1, 2), r2)
            rep(f"[{mode}] image features after Linear(65536,128)", o, r3)
            # the same Linear fed with the ORACLE's conv2 output (isolates the GEMM)
            a = r2.float().permute(0, 2, 3, 1).reshape(n, 65536).contiguous().to(dev)
            a_hi, a_lo = ops.cast16(a, fmt, want_lo=split)
            o2, _ = ops.gemm_bf16(a_hi, 65536, wfc, 128, bias=fc.bias, act="relu", split_k=8, fmt=fmt, a_lo=a_lo)
            rep(f"[{mode}] Linear(65536,128) alone, exact input", o2, r3)
            if split:
                # is the rest weight rounding, or the tensor core's accumulation over a long K?  more split-K partials
                # (summed in fp32 by the finish kernel) shorten every accumulation chain; split weights remove their rounding
                for sk in (1, 8, 32, 128):
                    o3, _ = ops.gemm_bf16(a_hi, 65536, wfc, 128, bias=fc.bias, act="relu", split_k=sk, fmt=fmt, a_lo=a_lo)
                    rep(f"[{mode}]   ... split_k = {sk}", o3, r3)
                w_hi, w_lo = ops.cast16(wfc.float(), fmt, want_lo=False)[0], None
                wexact = ops.fc_weight_to_hwc_bf16(fc.weight.detach(), 64, 1024, fmt)
                w32 = torch.empty((128, 65536), device=dev)
                w32.copy_(fc.weight.detach().view(128, 64, 1024).permute(0, 2, 1).reshape(128, 65536))
                wh, wl = ops.cast16(w32, fmt, want_lo=True)
                for sk in (8, 128):
                    o4, _ = ops.gemm_bf16(a_hi, 65536, wh, 128, bias=fc.bias, act="relu", split_k=sk, fmt=fmt, a_lo=a_lo, w_lo=wl)
                    rep(f"[{mode}]   ... split weights too (x3), split_k = {sk}", o4, r3)
        want = torch.cat([model(fp[i:i + 256], img[i:i + 256]).reshape(-1) for i in range(0, fp.shape[0], 256)])
        for mode in ("strict", "fp16", "bf16"):
            ours.set_precision(mode)
            got = ours.predict_batches(fp.to(dev), img.to(dev), 256).cpu()
            print(f"[{mode}] scores: max |d| {float((got - want).abs().max()):.3e} mean {float((got - want).abs().mean()):.3e}")
        # strict image branch + fp32 everything else, and the reverse
        ours.set_precision("strict")
        im_strict = ours._image_branch(img.to(dev))
        ours.set_precision("fp32")
        im_f32 = ours._image_branch(img.to(dev))
        print(f"image branch strict vs fp32 kernels: max {float((im_strict - im_f32).abs().max()):.3e}")
        keep = ours._image_branch
        for name, im in (("strict image branch, fp32 rest", im_strict), ("fp32 image branch, strict rest", im_f32)):
            outs = []
            for i in range(0, fp.shape[0], 256):
                ours._image_branch = lambda image, i=i, im=im: im[i:i + 256]
                ours.set_precision("fp32" if name.startswith("strict") else "strict")
                ours.use_cuda_graphs = False
                outs.append(ours(fp[i:i + 256].to(dev), img[i:i + 256].to(dev)).reshape(-1).cpu())
            got = torch.cat(outs)
            print(f"{name}: max |d| {float((got - want).abs().max()):.3e} mean {float((got - want).abs().mean()):.3e}")
        ours._image_branch = keep


if __name__ == "__main__":
    main()
