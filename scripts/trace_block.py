"""Phase timeline of the fused CFT-block kernel from its own clock samples (cft_debug_block_trace).
python scripts/trace_block.py --d 256 --batch 32"""
import argparse, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cft = importlib.import_module("multispectral-object-detection_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--d", type=int, default=256)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--cluster", type=int, default=0)
args = ap.parse_args()
dev = "cuda"
d, B = args.d, args.batch
g = cft.modules.GPT(d).eval().to(dev)
w = g._weights(torch.device(dev))["stack"]
tok = torch.randn(B, 128, d, device=dev)
L = w["layers"]
lib = cft._lib.lib()
for _ in range(3):
    cft.ops.gpt_block(tok, w, g.h, cluster=args.cluster)
torch.cuda.synchronize()
buf = torch.zeros(1024 * L * 48, dtype=torch.int64, device=dev)
lib.cft_debug_block_trace(buf.data_ptr())
cft.ops.gpt_block(tok, w, g.h, cluster=args.cluster)
torch.cuda.synchronize()
lib.cft_debug_block_trace(None)
t = buf.view(1024, L, 48).cpu()
used = (t[:, 1, 0] != 0).nonzero().flatten()
t = t[used].double()
names = ["layer start", "QKV acc ready", "QKV drained->tiles", "attention done", "#A passed", "out acc ready",
         "out epilogue done", "LN2 stats barrier (#B)", "LN2 published (#C)", "up acc ready", "up epilogue done",
         "#D passed", "down acc ready", "down epilogue done", "LN stats barrier (#E)", "LN published (#F)"]
print(f"d={d} batch={B} CTAs traced {len(used)}; mean cycles per phase over layers 1..{L-1} (compute warp 0)")
lay = t[:, 1:, :]
tot = (lay[:, :, 15] - lay[:, :, 0]).mean()
prev = lay[:, :, 0]
for s in range(1, 16):
    dt = (lay[:, :, s] - prev)
    print(f"  {names[s]:28s} {dt.mean():9.0f}  (min {dt.min():7.0f} max {dt.max():7.0f})")
    prev = lay[:, :, s]
print(f"  layer total {tot:.0f} cycles; whole kernel per CTA {(t[:, L-1, 15] - t[:, 0, 0]).mean():.0f}")
print("MMA issuer, per GEMM pass (cycles relative to the compute warp's layer start): start | first operands landed | last MMA issued")
for i in range(10):
    a = lay[:, :, 16 + 3 * i: 19 + 3 * i]
    if (a == 0).all():
        break
    rel = a - lay[:, :, 0:1]
    print(f"  pass {i}: {rel[:, :, 0].mean():9.0f} {rel[:, :, 1].mean():9.0f} {rel[:, :, 2].mean():9.0f}   issue span {(a[:, :, 2] - a[:, :, 1]).mean():7.0f}")
