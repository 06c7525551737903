#!/usr/bin/env python
"""Lean A/B timer: one process = one setting of the library's environment switches (they are read once at load).
Builds the bench model (yolov5l-x3 FLIR, batch 32 @ 640x640), captures the forward into a CUDA graph and times K
replays with CUDA events (same method as bench.py's `value`, without its other legs).  `--nms` also times cft_nms on a
batch-32 z.  Prints one JSON line."""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--tag", default="")
ap.add_argument("--nms", action="store_true")
args = ap.parse_args()
pkg = importlib.import_module("multispectral-object-detection_b200")
dev = torch.device("cuda", 0)
cfg = pkg.named_config("yolov5l_fusion_transformerx3_FLIR_aligned")
torch.manual_seed(0)
model = pkg.Model(cfg).eval()
g = torch.Generator().manual_seed(2)
with torch.no_grad():
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
model = model.to(dev)
B = args.batch
eng = pkg.ForwardEngine(model, B, 640, 640, device=dev, slots=1)
x = torch.randint(0, 256, (B, 6, 640, 640), dtype=torch.uint8, generator=torch.Generator().manual_seed(1))
eng.x_dev[0].copy_(x.to(dev))
for _ in range(5):
    eng.run_resident(0)
torch.cuda.synchronize()
best = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(eng.compute)
    for _ in range(args.steps):
        eng.run_resident(0)
    e1.record(eng.compute)
    torch.cuda.synchronize()
    best.append(e0.elapsed_time(e1) / args.steps)
flags = {k: v for k, v in os.environ.items() if k.startswith("CFT_")}
line = {"tag": args.tag, "flags": flags, "batch": B, "ms_per_step": [round(v, 4) for v in best],
        "pairs_per_s": round(B / (min(best) / 1e3), 1), "launches": eng.launches_per_forward,
        "z_checksum": float(eng.z_dev[0].double().abs().sum())}
if args.nms:
    nms = importlib.import_module("multispectral-object-detection_b200.nms")
    z = eng.z_dev[0]
    out = torch.zeros(B, 300, 6, device=dev)
    cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    ws = torch.empty(B * z.shape[1], dtype=torch.int64, device=dev)
    res = {}
    for name, pred, kw in (("model_z_conf0.25", z, {}), ("model_z_conf0.001", z, {"conf_thres": 0.001}),):
        for _ in range(2):
            nms.nms_batched(pred, out=out, counts=cnt, workspace=ws, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            nms.nms_batched(pred, out=out, counts=cnt, workspace=ws, **kw)
        e1.record()
        torch.cuda.synchronize()
        res[name] = {"us": round(e0.elapsed_time(e1) * 100, 1), "kept_mean": float(cnt.float().mean())}
    from oracle import nms_oracle as N      # input generator only
    p = N.make_predictions(B, 25200, 3, seed=41).to(dev)
    for _ in range(2):
        nms.nms_batched(p, out=out, counts=cnt, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        nms.nms_batched(p, out=out, counts=cnt, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    res["clustered_32x25200"] = {"us": round(e0.elapsed_time(e1) * 100, 1), "kept_mean": float(cnt.float().mean())}
    line["nms"] = res
print(json.dumps(line), flush=True)
