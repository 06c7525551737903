"""One warm launch of cft_nms on a batch-32 x 25200-row z (clustered synthetic boxes and the uniform case) for ncu."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nms = importlib.import_module("multispectral-object-detection_b200.nms")
from oracle import nms_oracle as N   # input generator only

dev = "cuda"
p = N.make_predictions(32, 25200, 3, seed=41).to(dev)
out = torch.zeros(32, 300, 6, device=dev)
cnt = torch.zeros(32, dtype=torch.int32, device=dev)
ws = torch.empty(32 * 25200, dtype=torch.int64, device=dev)
for _ in range(2):
    nms.nms_batched(p, out=out, counts=cnt, workspace=ws)
torch.cuda.synchronize()
print("kept", cnt.float().mean().item())
