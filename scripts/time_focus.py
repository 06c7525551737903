"""Times the fused Focus kernel alone (B=32, 640x640 uint8, Cout=64) and the gather+conv path beside it."""
import importlib, os, sys, math
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("multispectral-object-detection_b200")
ops = pkg.ops
torch.manual_seed(0)
w = torch.randn(64, 12, 3, 3) / math.sqrt(108); b = torch.randn(64) * 0.5
x6 = torch.randint(0, 256, (32, 6, 640, 640), dtype=torch.uint8, device="cuda")
wf, bf = ops.pack_focus_weight(w, b, None, device="cuda")
wp, bp = ops.pack_conv_weight(w, b, None, cin_pad_to=16, device="cuda")
out = ops.empty_nhwc(32, 64, 320, 320, "cuda")
def fused(): ops.focus_conv(x6[:, :3], wf, bf, 64, 1, out=out)
def two(): ops.conv2d(ops.focus_gather(x6[:, :3]), wp, bp, 3, 1, 1, out=out, cout=64, cin=16)
for name, fn in (("fused", fused), ("gather+conv", two)):
    g = torch.cuda.CUDAGraph()
    fn(); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(10): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"{name:12s} {us:7.1f} us per call   (out 419 MB + in 39 MB -> {458e6 / us / 1e6:.2f} TB/s; {2*32*320*320*64*108/us/1e6:.0f} TFLOP/s useful)")
