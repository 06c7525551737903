#!/usr/bin/env python
"""Experiment: several batch-32 forwards in flight.  Each in-flight batch has its own CUDA graph (own activation pool) and
its own stream; graphs of different batches overlap on the GPU (one batch's latency-bound CFT chains and kernel tails
leave SMs that the other batch's convolutions use).  Prints ms per batch-32 step for 1, 2, 3 batches in flight, and for
two half batches (2 x 16) as a same-work alternative.  Same timing method as bench.py's `value`."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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


def run(batch, inflight, steps=24):
    engs = [pkg.ForwardEngine(model, batch, 640, 640, device=dev, slots=1) for _ in range(inflight)]
    x = torch.randint(0, 256, (batch, 6, 640, 640), dtype=torch.uint8, generator=torch.Generator().manual_seed(1)).to(dev)
    for e in engs:
        e.x_dev[0].copy_(x)
    torch.cuda.synchronize()
    for _ in range(3):
        for e in engs:
            e.run_resident(0)
    torch.cuda.synchronize()
    main = torch.cuda.current_stream()
    res = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for e in engs:
            e.compute.wait_event(e0)
        for i in range(steps):
            engs[i % inflight].run_resident(0)
        for e in engs:
            main.wait_stream(e.compute)
        e1.record(main)
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / steps)
    z = [float(e.z_dev[0].double().abs().sum()) for e in engs]
    del engs
    torch.cuda.empty_cache()
    return res, z


out = {}
for batch, inflight in ((32, 1), (32, 2), (32, 3), (16, 2), (16, 4)):
    ms, z = run(batch, inflight)
    out[f"b{batch}x{inflight}"] = {"ms_per_graph": [round(v, 3) for v in ms], "pairs_per_s": round(batch / (min(ms) / 1e3), 1),
                                   "z_checksums_equal": len(set(z)) == 1}
    print(json.dumps({f"b{batch}x{inflight}": out[f"b{batch}x{inflight}"]}), flush=True)
