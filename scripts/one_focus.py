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
out = ops.empty_nhwc(32, 64, 320, 320, "cuda")
for _ in range(3):
    ops.focus_conv(x6[:, :3], wf, bf, 64, 1, out=out)
torch.cuda.synchronize()
