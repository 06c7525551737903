"""Time one CFT / GPT block (tokeniser excluded): fused one-launch stack vs the per-op path, CUDA events, L2 flushed
between repetitions.  python scripts/time_block.py [--batch 32]"""
import argparse, importlib, json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cft = importlib.import_module("multispectral-object-detection_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--dims", type=int, nargs="*", default=[256, 512, 1024])
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] \
    if os.path.isfile(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 1385.4


def timeit(fn, do_flush=True):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(args.reps):
        if do_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


for d in args.dims:
    B = args.batch
    g = cft.modules.GPT(d).eval().to(dev)
    w = g._weights(torch.device(dev))
    tok = torch.randn(B, 128, d, device=dev)
    gflop = B * (24576 * d * d + 524288 * d) / 1e9
    ops = cft.ops

    def per_op():
        x2d = tok.view(B * 128, d)
        for L in w["layers"]:
            y = ops.layernorm(x2d, *L["ln1"])
            qkv = ops.gemm(y, L["qkv"][0], L["qkv"][1])
            att = ops.attention(qkv, B, 128, d, g.h)
            x2d = ops.gemm(att, L["out"][0], L["out"][1], residual=x2d, out_dtype=torch.float32)
            y = ops.layernorm(x2d, *L["ln2"])
            hid = ops.gemm(y, L["up"][0], L["up"][1], act=ops.ACT_GELU)
            x2d = ops.gemm(hid, L["down"][0], L["down"][1], residual=x2d, out_dtype=torch.float32)
        return ops.layernorm(x2d, *w["lnf"], out_dtype=torch.float32)

    row = {"d": d, "batch": B, "gflop": round(gflop, 1)}
    # per-op path under a CUDA graph (how the model runs it)
    gph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        per_op()
        torch.cuda.synchronize()
        with torch.cuda.graph(gph, stream=s):
            per_op()
    ms = timeit(gph.replay)
    row["per_op_graph_ms"] = round(ms, 4); row["per_op_frac"] = round(gflop / ms / peak, 4)
    if ops.gpt_block_supported(B, d, g.h, 128):
        for c in [0]:
            try:
                ms = timeit(lambda: ops.gpt_block(tok, w["stack"], g.h, cluster=c))
                row[f"fused_c{c}_ms"] = round(ms, 4); row[f"fused_c{c}_frac"] = round(gflop / ms / peak, 4)
                row[f"fused_c{c}_warmL2_ms"] = round(timeit(lambda: ops.gpt_block(tok, w["stack"], g.h, cluster=c), False), 4)
            except Exception as e:          # noqa
                row[f"fused_c{c}"] = str(e)[:80]
    print(json.dumps(row), flush=True)
