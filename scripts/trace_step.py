"""Timeline of one CUDA-graph replay of the headline forward: every conv/GEMM launch reports its own
{first CTA start, last CTA end} (%globaltimer), so in-kernel time, gaps and stream overlap are measured in the
real steady state (graph + PDL + two streams), not under per-launch events.  python scripts/trace_step.py [batch]"""
import collections
import ctypes
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
pkg = importlib.import_module("multispectral-object-detection_b200")
ops, L = pkg.ops, pkg._lib
B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 32
one_stream = "--one-stream" in sys.argv
model = pkg.Model(pkg.named_config("yolov5l_fusion_transformerx3_FLIR_aligned")).eval().cuda()
model.two_streams = not one_stream
x6 = torch.randint(0, 256, (B, 6, 640, 640), dtype=torch.uint8, device="cuda")
names = []
orig_conv, orig_gemm = ops.conv2d, ops.gemm


def conv_wrap(x, w, bias, k, stride, act, out=None, residual=None, cout=None, cin=None, impl="tcgen05", kw=0, chain=None,
              skip_out=False):
    b, c, h, wd = x.shape
    names.append(f"conv {cin or c:4d}->{cout or w.shape[0]:4d} k{k}s{stride} {h}x{wd}" + (" +res" if residual is not None else "")
                 + (" +1x1" if chain is not None else ""))
    return orig_conv(x, w, bias, k, stride, act, out=out, residual=residual, cout=cout, cin=cin, impl=impl, kw=kw, chain=chain,
                     skip_out=skip_out)


def gemm_wrap(a, w, bias, act=0, out=None, residual=None, out_dtype=torch.bfloat16, n=None, impl="tcgen05"):
    names.append(f"gemm M{a.shape[0]} K{a.shape[1]:4d} N{n or w.shape[0]:4d}" + (" +res" if residual is not None else ""))
    return orig_gemm(a, w, bias, act=act, out=out, residual=residual, out_dtype=out_dtype, n=n, impl=impl)


with torch.no_grad():
    for _ in range(2):
        model(x6[:, :3], x6[:, 3:])
torch.cuda.synchronize()
MAXL = 512
buf = torch.zeros(2 * MAXL, dtype=torch.int64, device="cuda")
ops.conv2d, ops.gemm = conv_wrap, gemm_wrap
stream = torch.cuda.Stream()
g = torch.cuda.CUDAGraph()
L.check(L.lib().cft_debug_conv_spans(ctypes.c_void_p(buf.data_ptr()), MAXL))
with torch.no_grad(), torch.cuda.graph(g, stream=stream):
    z = model(x6[:, :3], x6[:, 3:])
L.check(L.lib().cft_debug_conv_spans(ctypes.c_void_p(0), 0))
n = len(names)
init = torch.zeros(2 * MAXL, dtype=torch.int64)
init[0::2] = torch.iinfo(torch.int64).max
init = init.cuda()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(5):
    buf.copy_(init)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        e0.record(stream)
        g.replay()
        e1.record(stream)
    torch.cuda.synchronize()
t = buf.cpu().view(-1, 2)[:n]
t0 = int(t[:, 0].min())
start = (t[:, 0] - t0).double() / 1e3
end = (t[:, 1] - t0).double() / 1e3
dur = end - start
print(f"batch {B}, {'one stream' if one_stream else 'two streams'}: graph replay {e0.elapsed_time(e1) * 1e3:.0f} us; {n} conv/gemm launches, "
      f"first start -> last end {float(end.max()):.0f} us, sum of in-kernel spans {float(dur.sum()):.0f} us")
# union of busy intervals (any conv kernel resident)
iv = sorted(zip(start.tolist(), end.tolist()))
busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
for s_, e_ in iv[1:]:
    if s_ > cur_e:
        busy += cur_e - cur_s
        cur_s, cur_e = s_, e_
    else:
        cur_e = max(cur_e, e_)
busy += cur_e - cur_s
print(f"time with at least one conv kernel resident: {busy:.0f} us; without any: {float(end.max()) - busy:.0f} us")
agg = collections.OrderedDict()
for i, nm in enumerate(names):
    a = agg.setdefault(nm, [0, 0.0])
    a[0] += 1
    a[1] += float(dur[i])
import re


def ideal_us(nm):
    """Roofline time of one launch of this shape: max(FLOPs / sustained bf16 peak, compulsory bytes / HBM peak)."""
    m = re.match(r"conv\s+(\d+)->\s*(\d+) k(\d)s(\d) (\d+)x(\d+)", nm)
    if m:
        cin, cout, k, st, h, w = map(int, m.groups())
        ho, wo = (h + st - 1) // st, (w + st - 1) // st
        fl = 2.0 * B * ho * wo * cout * cin * k * k
        by = 2.0 * (B * h * w * cin + B * ho * wo * cout * (2 if "+res" in nm else 1) + cout * cin * k * k)
        if "+1x1" in nm:                     # chained cout -> cout 1x1: its FLOPs, its output and its weights
            fl += 2.0 * B * ho * wo * cout * cout
            by += 2.0 * (B * ho * wo * cout + cout * cout)
    else:
        m = re.match(r"gemm M(\d+) K\s*(\d+) N\s*(\d+)", nm)
        mm, kk, nn = map(int, m.groups())
        fl = 2.0 * mm * kk * nn
        by = 2.0 * (mm * kk + mm * nn + kk * nn) + (8.0 * mm * nn if "+res" in nm else 0.0)
    return max(fl / 1385.4e12, by / 6584.8e9) * 1e6, fl, by


print(f"{'shape':44s} {'n':>3s} {'us tot':>8s} {'us each':>8s} {'roofline':>9s} {'of roof':>8s} {'lost us':>8s}  (roofline = max(FLOPs / 1385 TF, bytes / 6585 GB/s))")
tot_lost = 0.0
for nm, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    idl, fl, by = ideal_us(nm)
    tot_lost += us - c * idl
    print(f"{nm:44s} {c:3d} {us:8.0f} {us / c:8.1f} {idl:9.1f} {100 * c * idl / us:7.0f}% {us - c * idl:8.0f}")
print(f"sum of (span - roofline) over all launches: {tot_lost:.0f} us")
if "--list" in sys.argv:
    for i, nm in enumerate(names):
        print(f"{i:3d} {nm:44s} start {float(start[i]):8.1f} dur {float(dur[i]):6.1f}")
