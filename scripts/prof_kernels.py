"""One kernel of the path per invocation, at its yolov5l-x3 batch-32 @640 shape, launched 3 times (2 warm + 1) -- the
target of the per-kernel `ncu --set full` captures (scripts/gpu_ncu_all.sh; summaries in profiles/r02_ncu_*.txt).
python scripts/prof_kernels.py <case>      cases: see CASES"""
import importlib
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("multispectral-object-detection_b200")
ops = pkg.ops
DEV, B = "cuda", 32


def nhwc(*shape):
    return torch.randn(*shape, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def conv_case(cin, cout, h, w, k, s, res=False, b=B):
    x = nhwc(b, cin, h, w)
    wp, bp = ops.pack_conv_weight(torch.randn(cout, cin, k, k) / math.sqrt(cin * k * k), torch.zeros(cout), None, device=DEV)
    r = nhwc(b, cout, h // s, w // s) if res else None
    y = ops.conv2d(x, wp, bp, k, s, 1, cout=cout, residual=r)
    return lambda: ops.conv2d(x, wp, bp, k, s, 1, out=y, cout=cout, residual=r)


def chain_case(c, h, w, b=B):
    """3x3 c -> c (+ residual) with the next Bottleneck's 1x1 chained in its epilogue (chain mode)."""
    x, r = nhwc(b, c, h, w), nhwc(b, c, h, w)
    wp, bp = ops.pack_conv_weight(torch.randn(c, c, 3, 3) / math.sqrt(9 * c), torch.zeros(c), None, device=DEV)
    w2, b2 = ops.pack_conv_weight(torch.randn(c, c, 1, 1) / math.sqrt(c), torch.zeros(c), None, device=DEV)
    y, y2 = nhwc(b, c, h, w), nhwc(b, c, h, w)
    return lambda: ops.conv2d(x, wp, bp, 3, 1, 1, out=y, cout=c, residual=r, chain=(w2, b2, 1, y2))


def gemm_case(m, kdim, n, act=0, res32=False):
    a = torch.randn(m, kdim, device=DEV).to(torch.bfloat16)
    wp, bp = ops.pack_linear_weight(torch.randn(n, kdim) / math.sqrt(kdim), torch.zeros(n), device=DEV)
    r = torch.randn(m, n, device=DEV) if res32 else None
    return lambda: ops.gemm(a, wp, bp, act=act, residual=r, out_dtype=torch.float32 if res32 else torch.bfloat16)


def gpt_block_case(d):
    g = pkg.modules.GPT(d).eval().to(DEV)
    w = g._weights(torch.device(DEV))["stack"]
    tok = torch.randn(B, 128, d, device=DEV)
    return lambda: ops.gpt_block(tok, w, g.h)


CASES = {
    # the tcgen05 implicit-GEMM kernel, one launch per shape class (SURVEY.md section 8d catalogue)
    "conv_c3_p3_3x3_128_rowreuse_pairs": lambda: conv_case(128, 128, 80, 80, 3, 1, res=True),
    "conv_c3_p4_3x3_256_pairs_n256": lambda: conv_case(256, 256, 40, 40, 3, 1, res=True),
    "conv_c3_p5_3x3_512_tiles_4x4x8": lambda: conv_case(512, 512, 20, 20, 3, 1, res=True),
    "conv_c3_p3_3x3_128_chain_1x1": lambda: chain_case(128, 80, 80),
    "conv_c3_p2_3x3_64_resident_weights": lambda: conv_case(64, 64, 160, 160, 3, 1, res=True),
    "conv_c3_p3_1x1_128_flat_hbm": lambda: conv_case(128, 128, 80, 80, 1, 1),
    "conv_down_p4_3x3s2_256_512": lambda: conv_case(256, 512, 80, 80, 3, 2),
    "conv_gpt_p5_up_gemm_1024_4096": lambda: gemm_case(4096, 1024, 4096, act=2),
    "conv_gpt_p5_out_gemm_f32_res": lambda: gemm_case(4096, 1024, 1024, res32=True),
    "focus_fused_u8": lambda: (lambda img, wf, bf: (lambda: ops.focus_conv(img, wf, bf, 64, 1)))(
        torch.randint(0, 256, (B, 3, 640, 640), dtype=torch.uint8, device=DEV),
        *ops.pack_focus_weight(torch.randn(64, 12, 3, 3) / 10, torch.zeros(64), None, device=DEV)),
    "attention_d1024": lambda: (lambda qkv: (lambda: ops.attention(qkv, B, 128, 1024, 8)))(
        torch.randn(B * 128, 3072, device=DEV).to(torch.bfloat16)),
    "attention_d512": lambda: (lambda qkv: (lambda: ops.attention(qkv, B, 128, 512, 8)))(
        torch.randn(B * 128, 1536, device=DEV).to(torch.bfloat16)),
    "layernorm_d1024": lambda: (lambda x, g, b: (lambda: ops.layernorm(x, g, b, 1e-5)))(
        torch.randn(B * 128, 1024, device=DEV), torch.ones(1024, device=DEV), torch.zeros(1024, device=DEV)),
    "pool_tokens_p3": lambda: (lambda r, i, p: (lambda: ops.gpt_pool_tokens(r, i, p, 8, 8)))(
        nhwc(B, 256, 80, 80), nhwc(B, 256, 80, 80), torch.zeros(1, 128, 256, device=DEV)),
    "unpool_p3_add2_add": lambda: (lambda t, r, i: (lambda: ops.gpt_unpool(t, 80, 80, 8, 8, x_rgb=r, x_ir=i, want_sum=True)))(
        torch.randn(B, 128, 256, device=DEV), nhwc(B, 256, 80, 80), nhwc(B, 256, 80, 80)),
    "spp_maxpool_cascade": lambda: (lambda cat: (lambda: ops.maxpool_cascade3(cat[:, :512], cat, [512, 1024, 1536], [5, 5, 5])))(
        nhwc(B, 2048, 20, 20)),
    "upsample2x_p4": lambda: (lambda x: (lambda: ops.upsample2x(x)))(nhwc(B, 256, 40, 40)),
    "detect_decode_p3": lambda: (lambda h, a, z: (lambda: ops.detect_decode(h, B, 80, 80, 3, 8, 8.0, a, z, 0)))(
        torch.randn(B * 6400, 24, device=DEV), torch.tensor([10., 13, 16, 30, 33, 23], device=DEV),
        torch.zeros(B, 25200, 8, device=DEV)),
    "gpt_block_d256": lambda: gpt_block_case(256),
    "gpt_block_d512": lambda: gpt_block_case(512),
}

if __name__ == "__main__":
    case = sys.argv[1]
    fn = CASES[case]()
    with torch.no_grad():
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    print("ran", case)
