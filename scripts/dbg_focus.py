import importlib, os, sys, math
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("multispectral-object-detection_b200")
ops = pkg.ops
torch.manual_seed(0)
cout = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w = torch.randn(cout, 12, 3, 3) / math.sqrt(108)
b = torch.randn(cout) * 0.5
img = torch.randint(0, 256, (2, 3, 64, 96), dtype=torch.uint8).cuda()
wf, bf = ops.pack_focus_weight(w, b, None, device="cuda")
y = ops.focus_conv(img, wf, bf, cout, 0)
torch.cuda.synchronize()
x = img.float() / 255
s2d = torch.cat([x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]], 1)
ref = torch.nn.functional.conv2d(s2d, w.cuda(), b.cuda(), padding=1)
d = (y.float() - ref).abs()
print("max abs diff", d.max().item(), "ref max", ref.abs().max().item())
