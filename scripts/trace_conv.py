"""In-kernel timeline of one cft_conv2d launch (debug): per-CTA clock samples written by the kernel itself.
python scripts/trace_conv.py <shape name from prof_shapes.SHAPES> [batch]"""
import ctypes
import importlib
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
pkg = importlib.import_module("multispectral-object-detection_b200")
ops, L = pkg.ops, pkg._lib
from prof_shapes import SHAPES  # noqa: E402

SLOTS = 64


def main():
    name = sys.argv[1]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    res = "--res" in sys.argv
    _, cin, cout, h, w, k, s = [r for r in SHAPES if r[0] == name][0]
    b = 1 if h == 1 else B
    x = torch.randn(b, cin, h, w, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wt = torch.randn(cout, cin, k, k) / math.sqrt(cin * k * k)
    wp, bp = ops.pack_conv_weight(wt, torch.zeros(cout), None, device="cuda")
    r = torch.randn(b, cout, h // s, w // s, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last) if res else None
    for _ in range(3):
        ops.conv2d(x, wp, bp, k, s, 1, cout=cout, residual=r)
    torch.cuda.synchronize()
    buf = torch.zeros(4096 * SLOTS, dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for cold in (0, 1):
        buf.zero_()
        if cold:
            flush.zero_()
        torch.cuda.synchronize()
        L.check(L.lib().cft_debug_conv_trace(ctypes.c_void_p(buf.data_ptr())))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv2d(x, wp, bp, k, s, 1, cout=cout, residual=r)
        e1.record()
        torch.cuda.synchronize()
        L.check(L.lib().cft_debug_conv_trace(ctypes.c_void_p(0)))
        t = buf.view(-1, SLOTS).cpu()
        n = int((t[:, 1] != 0).sum())
        t = t[:n]
        print(f"== {name} B={b} {'L2 flushed' if cold else 'warm'}: {n} CTAs, event time {e0.elapsed_time(e1) * 1e3:.1f} us")
        g0 = int(t[:, 0].min())
        print(f"   CTA start skew (globaltimer): max {int(t[:, 0].max()) - g0} ns")
        wall_ns = (t[:, SLOTS - 1] - t[:, 0]).double()
        cyc = (t[:, 7] - t[:, 1]).double()
        khz = float((cyc / wall_ns).median()) * 1e3          # SM cycles per us, calibrated with %globaltimer
        print(f"   kernel span (globaltimer, first start -> last exit): {(int(t[:, SLOTS - 1].max()) - g0) / 1e3:.1f} us; SM clock {khz:.0f} MHz")
        def col(i):
            return (t[:, i] - t[:, 1]).double() / khz
        for lab, i in (("setup done", 2), ("PDL wait done", 3), ("first stage landed", 4), ("last MMA committed", 5),
                       ("epilogue drained", 6), ("exit", 7)):
            c = col(i)[t[:, i] != 0]
            if len(c):
                print(f"   {lab:22s} min {c.min():7.2f}  median {c.median():7.2f}  max {c.max():7.2f} us   ({len(c)} CTAs)")
        tot = (t[:, 7] - t[:, 1]).double()
        for lab, i in (("MMA warp waiting for operands", SLOTS - 2), ("MMA warp waiting for a free accumulator", SLOTS - 3),
                       ("producer waiting for a free ring slot", SLOTS - 4)):
            c = t[:, i].double()
            m = c > 0
            if m.any():
                print(f"   {lab:42s} median {float((c[m] / tot[m]).median()) * 100:5.1f} % of the CTA's lifetime")
        for cta in (0, 1, 2, n // 2, n - 1):
            row = t[cta]
            ev = []
            for j in range((SLOTS - 12) // 2):
                if row[8 + 2 * j] or row[9 + 2 * j]:
                    ev.append(f"t{j}:{(int(row[8 + 2 * j]) - int(row[1])) / khz:.1f}-{(int(row[9 + 2 * j]) - int(row[1])) / khz:.1f}")
            print(f"   CTA {cta}: acc ready-released (us): " + " ".join(ev))


if __name__ == "__main__":
    main()
