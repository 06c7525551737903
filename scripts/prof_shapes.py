"""Runs the dominant conv/GEMM shapes of yolov5l-x3 @640, batch 32, once each after a warm-up (for ncu
--set full captures and for per-shape CUDA-event timing).  python scripts/prof_shapes.py [--time]"""
import importlib
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("multispectral-object-detection_b200")
ops = pkg.ops
DEV = "cuda"
B = 32
# name, Cin, Cout, H, W, k, s   (SURVEY.md §8d shape catalogue)
SHAPES = [
    ("c3_p3_3x3_128", 128, 128, 80, 80, 3, 1),
    ("c3_p4_3x3_256", 256, 256, 40, 40, 3, 1),
    ("c3_p5_3x3_512", 512, 512, 20, 20, 3, 1),
    ("c3_p2_3x3_64", 64, 64, 160, 160, 3, 1),
    ("down_p2_3x3s2_64_128", 64, 128, 320, 320, 3, 2),
    ("down_p4_3x3s2_256_512", 256, 512, 80, 80, 3, 2),
    ("c3_p3_1x1_128", 128, 128, 80, 80, 1, 1),
    ("c3_p2_1x1_64", 64, 64, 160, 160, 1, 1),
    ("c3_p4_cv12_512_512", 512, 512, 40, 40, 1, 1),
    ("spp_cv2_2048_1024", 2048, 1024, 20, 20, 1, 1),
    ("focus_16_64", 16, 64, 320, 320, 3, 1),
    ("focus_wide_64_64_3x1", 64, 64, 320, 320, 31, 1),
    ("gpt_p5_up_1024_4096", 1024, 4096, 1, 4096, 1, 1),
    ("gpt_p5_down_4096_1024", 4096, 1024, 1, 4096, 1, 1),
    ("gpt_p3_qkv_256_768", 256, 768, 1, 4096, 1, 1),
]


def main():
    timed = "--time" in sys.argv
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    act = 0 if "--noact" in sys.argv else 1
    peaks_t, peaks_b = 1385.4e12, 6584.8e9
    for name, cin, cout, h, w, k, s in SHAPES:
        if only and name not in only:
            continue
        b = 1 if h == 1 else B
        kw = 0
        if k == 31:
            k, kw = 3, 1
        kww = kw if kw else k
        x = torch.randn(b, cin, h, w, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        wt = torch.randn(cout, cin, k, kww) / math.sqrt(cin * k * kww)
        wp, bp = ops.pack_conv_weight(wt, torch.zeros(cout), None, device=DEV)
        y = ops.conv2d(x, wp, bp, k, s, act, cout=cout, kw=kw)            # warm-up
        torch.cuda.synchronize()
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
            ts = []
            for _ in range(5):
                flush.zero_()                                     # L2 flush between timed iterations
                e0.record()
                ops.conv2d(x, wp, bp, k, s, act, out=y, cout=cout, kw=kw)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            ho, wo = (h + s - 1) // s, (w + s - 1) // s
            flops = 2.0 * b * ho * wo * cout * cin * k * kww
            byts = 2.0 * (b * h * w * cin + b * ho * wo * cout + cout * cin * k * kww)
            print(f"{name:28s} {ms*1e3:9.1f} us  {flops/ms/1e9:8.1f} TFLOP/s ({flops/ms/1e9*1e12/peaks_t*100:5.1f}% tensor)  "
                  f"{byts/ms/1e6:8.1f} GB/s ({byts/ms/1e6*1e9/peaks_b*100:5.1f}% hbm)  AI {flops/byts:6.0f}", flush=True)
        else:
            ops.conv2d(x, wp, bp, k, s, act, out=y, cout=cout, kw=kw)
            torch.cuda.synchronize()


if __name__ == "__main__":
    main()
