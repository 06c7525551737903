#!/usr/bin/env python
"""Which role bounds the fused Focus kernel?  Times cft_focus_conv at the bench shape (batch 32, 640 x 640 uint8) for
Cout 32 / 64 / 128 with and without SiLU: the builders' work per tile does not depend on Cout, the epilogue's / the stores' does.
CUDA events around 20 launches after 5 warm-up launches; one JSON line per variant."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("multispectral-object-detection_b200")
ops = pkg.ops
B = 32
img = torch.randint(0, 256, (B, 3, 640, 640), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for cout in (32, 64, 128):
    wf, bf = ops.pack_focus_weight(torch.randn(cout, 12, 3, 3) / 10, torch.zeros(cout), None, device="cuda")
    for act in (1, 0):
        out = ops.focus_conv(img, wf, bf, cout, act)
        for _ in range(5):
            ops.focus_conv(img, wf, bf, cout, act, out=out)
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.focus_conv(img, wf, bf, cout, act, out=out)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        mb = (img.numel() + out.numel() * 2) / 1e6
        print(json.dumps({"cout": cout, "act": act, "us": round(us, 1), "MB": round(mb, 1), "GB_per_s": round(mb / us * 1e3, 0)}))
