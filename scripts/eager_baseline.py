#!/usr/bin/env python
"""The 'existing GPU path' baseline SURVEY.md section 8d asks for: the reference's forward as plain PyTorch eager ops
(cuDNN / cuBLAS through torch.nn.functional) on the SAME B200, half precision, channels_last, BN fused -- i.e. what
`test.py --half` / `detect_twostream.py` run on a GPU (test.py:66-68,107; detect_twostream.py:40-41,72).

/root/reference does not exist on the GPU box, so the op sequence comes from the oracle's restatement of the reference
forward (oracle/cft_oracle.py: identical torch.nn.functional calls, pinned bit-exact to the reference on CPU) with its
tensors moved to the GPU.  A reported baseline, never part of the product path; run by hand / by scripts/gpu_*.sh:

    python scripts/eager_baseline.py [--batch 32] [--steps 5] [--config yolov5l_fusion_transformerx3_FLIR_aligned]
"""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cft_oracle as O  # noqa: E402


def fuse_state(sd):
    """Model.fuse() (models/yolo_test.py:296-304, utils/torch_utils.py:181-201) on a state dict."""
    out = {}
    for k, v in sd.items():
        if ".bn." in k:
            continue
        if k.endswith("conv.weight") and k.replace("conv.weight", "bn.weight") in sd:
            p = k[:-len("conv.weight")]
            g, b = sd[p + "bn.weight"], sd[p + "bn.bias"]
            m, var = sd[p + "bn.running_mean"], sd[p + "bn.running_var"]
            scale = g / torch.sqrt(var + O.BN_EPS)
            out[k] = v * scale.view(-1, 1, 1, 1)
            out[p + "conv.bias"] = b - m * scale
        else:
            out[k] = v
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--config", default="yolov5l_fusion_transformerx3_FLIR_aligned")
    ap.add_argument("--size", type=int, nargs=2, default=[640, 640])
    args = ap.parse_args()
    pkg = importlib.import_module("multispectral-object-detection_b200")   # config dicts only
    cfg = pkg.named_config(args.config)
    dev = torch.device("cuda", 0)
    sd32 = fuse_state(O.init_state(cfg, seed=0))
    h, w = args.size
    flops = O.conv_linear_flops(cfg, h, w)
    for dtype, bench_flag in ((torch.bfloat16, False), (torch.float16, False), (torch.float16, True)):
        torch.backends.cudnn.benchmark = bench_flag
        sd = {}
        for k, v in sd32.items():
            if v.is_floating_point():
                v = v.to(dev, dtype)
                if v.dim() == 4:
                    v = v.contiguous(memory_format=torch.channels_last)
            else:
                v = v.to(dev)
            sd[k] = v
        g = torch.Generator().manual_seed(1)
        x = torch.rand(args.batch, 3, h, w, generator=g).to(dev, dtype).contiguous(memory_format=torch.channels_last)
        x2 = torch.rand(args.batch, 3, h, w, generator=g).to(dev, dtype).contiguous(memory_format=torch.channels_last)
        for _ in range(3 if bench_flag else 2):
            z, _ = O.forward(sd, cfg, x, x2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            z, _ = O.forward(sd, cfg, x, x2)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({"baseline": "pytorch eager (cuDNN/cuBLAS), BN fused, channels_last", "config": args.config,
                          "dtype": str(dtype).replace("torch.", ""), "cudnn_benchmark": bench_flag, "batch": args.batch,
                          "height": h, "width": w, "ms_per_step": round(ms, 3),
                          "pairs_per_s": round(args.batch / (ms / 1e3), 1),
                          "tflops": round(flops * args.batch / (ms / 1e3) / 1e12, 1),
                          "z_finite": bool(torch.isfinite(z.float()).all())}), flush=True)
        del sd, x, x2, z
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
