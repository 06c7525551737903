"""Where does the end-to-end (host->device->host) time go?  python scripts/e2e_debug.py"""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("multispectral-object-detection_b200")
B, H, W, K = 32, 640, 640, 20
model = pkg.Model(pkg.named_config("yolov5l_fusion_transformerx3_FLIR_aligned")).eval().cuda()
eng = pkg.ForwardEngine(model, B, H, W, device="cuda", slots=2)
host = torch.randint(0, 256, (B, 6, H, W), dtype=torch.uint8).pin_memory()
dev = torch.empty_like(host, device="cuda")
zh = eng.z_host[0]


def timeit(fn, n=K):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


print("replay slot0        %.3f ms" % timeit(lambda: eng.run_resident(0)))
print("replay slot1        %.3f ms" % timeit(lambda: eng.run_resident(1)))
alt = [0]
def both():
    eng.run_resident(alt[0]); alt[0] ^= 1
print("replay alternating  %.3f ms" % timeit(both))
print("H2D 78.6 MB         %.3f ms" % timeit(lambda: dev.copy_(host, non_blocking=True)))
print("D2H 25.8 MB         %.3f ms" % timeit(lambda: zh.copy_(eng.z_dev[0], non_blocking=True)))
def e2e():
    if len(eng._pending) == eng.slots:
        eng.collect()
    eng.submit(host)
t = timeit(e2e); eng.drain()
print("e2e pipelined       %.3f ms" % t)
def serial():
    eng.infer(host)
print("e2e blocking infer  %.3f ms" % timeit(serial, 10))
