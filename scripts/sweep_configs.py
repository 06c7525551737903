#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations (the bench line covers config 2 only):

  config 3  yolov5l_fusion_transformerx3 (LLVIP cfg, nc 1) @ 1024x1280, batch 8
  config 5  yolov5x_fusion_transformerx3 @ 640x640, batch sweep 1..128
  (+ the yolov5l FLIR batch sweep, to show where the GPT blocks stop being weight-bandwidth-bound)

Same method as bench.py's `value`: CUDA-graph replay on device-resident uint8 inputs, CUDA events around K replays after
W warm-up replays; one JSON line per point on stdout (`--out` also writes them to a file).  Algorithmic FLOPs per pair
from the oracle's FLOP model; tensor fraction against MEASURED_PEAKS.json's sustained bf16 peak.
"""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None, help="substring filter on the point name")
    args = ap.parse_args()
    pkg = importlib.import_module("multispectral-object-detection_b200")
    from oracle import cft_oracle as O           # FLOP model only
    sys.path.insert(0, ROOT)
    import bench
    peaks = bench.load_peaks()
    dev = torch.device("cuda", 0)
    points = [("config3_l_llvip_1024x1280", "yolov5l_fusion_transformerx3_llvip", 8, 1024, 1280)]
    points += [(f"config5_x_640_b{b}", "yolov5x_fusion_transformerx3_FLIR_aligned", b, 640, 640)
               for b in (1, 2, 4, 8, 16, 32, 64, 128)]
    points += [(f"l_flir_640_b{b}", "yolov5l_fusion_transformerx3_FLIR_aligned", b, 640, 640) for b in (1, 4, 8, 16, 64)]
    lines = []
    models = {}
    for name, cname, b, h, w in points:
        if args.only and args.only not in name:
            continue
        cfg = pkg.named_config(cname)
        if cname not in models:
            torch.manual_seed(0)
            m = pkg.Model(cfg).eval()
            g = torch.Generator().manual_seed(2)
            with torch.no_grad():
                for mod in m.modules():
                    if isinstance(mod, torch.nn.BatchNorm2d):
                        mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
                        mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
                        mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) + 0.5)
                        mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
            models = {cname: m.to(dev)}           # one model resident at a time
        model = models[cname]
        eng = pkg.ForwardEngine(model, b, h, w, device=dev, slots=1)
        x = torch.randint(0, 256, (b, 6, h, w), dtype=torch.uint8, generator=torch.Generator().manual_seed(1))
        eng.x_dev[0].copy_(x.to(dev))
        for _ in range(max(args.warmup, 3)):
            eng.run_resident(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(eng.compute)
        for _ in range(args.steps):
            eng.run_resident(0)
        e1.record(eng.compute)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        flops = O.conv_linear_flops(cfg, h, w)
        pairs_s = b / (ms / 1e3)
        line = {"point": name, "config": cname, "batch": b, "height": h, "width": w, "ms_per_step": round(ms, 4),
                "pairs_per_s": round(pairs_s, 2), "gflop_per_pair": round(flops / 1e9, 2),
                "tensor_frac_sustained": round(flops * pairs_s / (peaks["tflops_sustained"] * 1e12), 4),
                "launches_per_step": eng.launches_per_forward, "steps": args.steps, "launch": "cuda-graph replay"}
        print(json.dumps(line), flush=True)
        lines.append(line)
        del eng
        torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
