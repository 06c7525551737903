"""Per-shape time of every conv/GEMM launch in one forward of the headline workload (CUDA events around each
call, L2 not flushed: in-situ timing). python scripts/profile_layers.py [batch]"""
import collections
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("multispectral-object-detection_b200")
ops = pkg.ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
model = pkg.Model(pkg.named_config("yolov5l_fusion_transformerx3_FLIR_aligned")).eval().cuda()
model.two_streams = False      # per-launch events only add up when the walk is serialised on one stream
x6 = torch.randint(0, 256, (B, 6, 640, 640), dtype=torch.uint8, device="cuda")
records = []
orig_conv, orig_gemm = ops.conv2d, ops.gemm


def timed(fn, key_fn):
    def wrap(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        records.append((key_fn(*a, **k), e0, e1))
        return out
    return wrap


def conv_key(x, w, bias, k, stride, act, out=None, residual=None, cout=None, cin=None, impl="tcgen05", kw=0):
    b, c, h, wd = x.shape
    cin = cin or c
    cout = cout or w.shape[0]
    ho, wo = (h + stride - 1) // stride, (wd + stride - 1) // stride
    kww = kw if kw else k
    fl = 2.0 * b * ho * wo * cout * cin * k * kww
    by = 2.0 * (b * h * wd * cin + b * ho * wo * cout * (2 if residual is not None else 1) + cout * cin * k * kww)
    return (f"conv {cin:4d}->{cout:4d} k{k}x{kww}s{stride} {h}x{wd}" + (" +res" if residual is not None else ""), fl, by)


def gemm_key(a, w, bias, act=0, out=None, residual=None, out_dtype=torch.bfloat16, n=None, impl="tcgen05"):
    m, kd = a.shape
    n = n or w.shape[0]
    osz = 4 if (out_dtype == torch.float32 or (out is not None and out.dtype == torch.float32)) else 2
    return (f"gemm M{m} K{kd:4d} N{n:4d}" + (" f32+res" if residual is not None else ""), 2.0 * m * kd * n,
            2.0 * m * kd + osz * m * n * (2 if residual is not None else 1) + 2.0 * kd * n)


ops.conv2d = timed(orig_conv, conv_key)
ops.gemm = timed(orig_gemm, gemm_key)
with torch.no_grad():
    for it in range(3):
        records.clear()
        model(x6[:, :3], x6[:, 3:])
        torch.cuda.synchronize()
agg = collections.OrderedDict()
for (name, fl, by), e0, e1 in records:
    a = agg.setdefault(name, [0, 0.0, fl, by])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print(f"batch {B}: {len(records)} conv/gemm launches, {tot:.2f} ms total")
print(f"{'shape':44s} {'n':>3s} {'ms tot':>8s} {'us each':>8s} {'TF/s':>7s} {'%tens':>6s} {'GB/s':>7s} {'%hbm':>5s}")
for name, (n, ms, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    each = ms / n
    print(f"{name:44s} {n:3d} {ms:8.3f} {each*1e3:8.1f} {fl/each/1e9:7.1f} {fl/each/1e9/1385.4*100:6.1f} {by/each/1e6:7.0f} {by/each/1e6/6584.8*100:5.1f}")
