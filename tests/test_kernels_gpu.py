"""GPU parity tests of every C-ABI kernel against a plain fp32 PyTorch restatement of the same op
(the per-module half of the parity bar; the whole-model half is test_model_gpu.py vs the CPU oracle).

Tolerances (stated per SURVEY.md §8c): inputs/weights are bf16-rounded for BOTH sides, so a bf16
output may differ by its own rounding (rel 2^-8) plus fp32 accumulation-order noise:
|d| <= 1e-2 * max|ref| + 2e-2 * |ref| for bf16 outputs, 1e-3 relative for fp32 outputs.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale)


def nhwc(t):
    return t.to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def close_bf16(out, ref, what=""):
    out, ref = out.float(), ref.float()
    tol = 1e-2 * ref.abs().max() + 2e-2 * ref.abs()
    bad = ((out - ref).abs() > tol)
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} off, max|d|={float((out - ref).abs().max()):.4g} ref max {float(ref.abs().max()):.4g}"


def conv_ref(x, w, b, k, s, act, res=None):
    y = F.conv2d(x.float(), w.float(), b.float() if b is not None else None, stride=s, padding=k // 2)
    if act == 1:
        y = F.silu(y)
    elif act == 2:
        y = F.gelu(y)
    if res is not None:
        y = y + res.float()
    return y


CONV_CASES = [
    # B, Cin, Cout, H, W, k, s, act, residual
    (2, 64, 64, 16, 16, 1, 1, 1, False),
    (2, 64, 128, 32, 32, 3, 1, 1, True),
    (1, 128, 128, 20, 20, 3, 1, 1, True),       # 20x20: non-power-of-two spatial tile (P5)
    (2, 64, 128, 32, 32, 3, 2, 1, False),       # stride 2 through the parity maps
    (1, 256, 512, 40, 40, 3, 2, 1, False),      # n_blocks = 2
    (3, 32, 32, 24, 40, 1, 1, 1, False),        # Cin < 64 (yolov5s), OOB K fill
    (1, 80, 160, 16, 16, 3, 1, 1, True),        # yolov5x widths: K tail 80 = 64 + 16, N = 160
    (2, 160, 160, 20, 20, 3, 1, 1, True),       # yolov5x: K = 160 = 2 x 64 + 32 per tap
    (1, 16, 64, 64, 64, 3, 1, 1, False),        # Focus-style 16-channel input
    (1, 512, 24, 20, 20, 1, 1, 0, False),       # Detect-style N = 24
    (2, 1024, 1024, 8, 8, 1, 1, 1, False),      # 16 k-chunks, 4 n-blocks
    (1, 320, 640, 8, 12, 1, 1, 0, False),       # N = 640 -> 4 x 160
    (2, 32, 64, 24, 40, 3, 1, 1, True),         # 64 B operand rows (SWIZZLE_64B ring), residual via TMA
    (2, 16, 32, 32, 32, 3, 2, 1, False),        # 32 B operand rows (SWIZZLE_32B ring), stride 2
    (1, 64, 96, 20, 20, 3, 1, 1, True),         # 3 column chunks: uneven split over the 2 epilogue groups
    (32, 128, 128, 80, 80, 3, 1, 1, True),      # full-size C3 bottleneck conv (batch 32): many tiles per CTA
    (4, 128, 128, 40, 40, 3, 1, 1, True),       # batch-spanning tiles: 8 x 8 px x 2 images, residual through the same boxes
    (8, 64, 128, 20, 20, 3, 1, 1, False),       # 4 x 4 px x 8 images
    (3, 64, 64, 20, 20, 3, 1, 1, True),         # ragged: 3 images in tiles of up to 8 (out-of-range images clipped / zero-filled)
    (5, 64, 128, 40, 40, 3, 2, 1, False),       # stride 2 -> 20 x 20 output, 5 images over 4 x 4 x 8 tiles
    (6, 256, 256, 40, 40, 3, 1, 1, True),       # CTA pairs over batch-spanning tiles, N = 256
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_tcgen05_matches_torch_and_cuda_core_ref(case, cft):
    B, Cin, Cout, H, W, k, s, act, use_res = case
    ops = cft.ops
    x = nhwc(rnd(B, Cin, H, W, seed=1))
    w = rnd(Cout, Cin, k, k, seed=2, scale=1.0 / math.sqrt(Cin * k * k))
    b = rnd(Cout, seed=3, scale=0.5)
    wp, bp = ops.pack_conv_weight(w, b, None, device=DEV)
    Ho, Wo = (H + s - 1) // s, (W + s - 1) // s
    res = nhwc(rnd(B, Cout, Ho, Wo, seed=4)) if use_res else None
    y = ops.conv2d(x, wp, bp, k, s, act, residual=res, cout=Cout)
    y_ref_kernel = ops.conv2d(x, wp, bp, k, s, act, residual=res, cout=Cout, impl="ref")
    torch.cuda.synchronize()
    ref = conv_ref(x, wp[:Cout, :, :Cin].float().reshape(Cout, k, k, Cin).permute(0, 3, 1, 2), bp[:Cout], k, s, act, res)
    assert y.shape == (B, Cout, Ho, Wo) and y.is_contiguous(memory_format=torch.channels_last)
    close_bf16(y_ref_kernel, ref, "cuda-core ref kernel vs torch")
    close_bf16(y, ref, "tcgen05 vs torch")
    # the two kernels see identical bf16 operands: only accumulation order differs
    assert (y.float() - y_ref_kernel.float()).abs().max() <= 2e-2 * ref.abs().max()


def test_conv_writes_into_channel_slice_and_reads_slice(cft):
    """Concat fusion: input is a channel slice of a wider buffer, output goes into a slice of another."""
    ops = cft.ops
    B, H, W = 2, 16, 24
    wide = nhwc(rnd(B, 192, H, W, seed=5))
    x = wide[:, 64:128]
    w = rnd(64, 64, 3, 3, seed=6, scale=1 / 24.0)
    wp, bp = ops.pack_conv_weight(w, None, None, device=DEV)
    dst = torch.full((B, 160, H, W), 7.0, device=DEV, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ops.conv2d(x, wp, bp, 3, 1, 1, out=dst[:, 32:96], cout=64)
    torch.cuda.synchronize()
    ref = conv_ref(x, wp.float().reshape(64, 3, 3, 64).permute(0, 3, 1, 2), None, 3, 1, 1)
    close_bf16(dst[:, 32:96], ref, "slice conv")
    assert (dst[:, :32] == 7).all() and (dst[:, 96:] == 7).all()      # neighbours untouched


@pytest.mark.parametrize("M,K,N,act,f32out", [(256, 256, 768, 0, False), (384, 512, 2048, 2, False),
                                               (256, 1024, 256, 0, True), (128, 160, 480, 0, False),
                                               (4096, 256, 256, 0, True)])
def test_gemm_linear(M, K, N, act, f32out, cft):
    """nn.Linear of the CFT blocks: out = act(a @ w^T + b) [+ residual] (fp32 residual stream)."""
    ops = cft.ops
    a = rnd(M, K, seed=7).to(DEV).to(torch.bfloat16)
    w = rnd(N, K, seed=8, scale=0.05)
    b = rnd(N, seed=9, scale=0.1)
    wp, bp = ops.pack_linear_weight(w, b, device=DEV)
    res = rnd(M, N, seed=10).to(DEV) if f32out else None
    out = ops.gemm(a, wp, bp, act=act, residual=res, out_dtype=torch.float32 if f32out else torch.bfloat16)
    torch.cuda.synchronize()
    ref = a.float() @ wp.view(N, K).float().t() + bp
    if act == 2:
        ref = F.gelu(ref)
    if res is not None:
        ref = ref + res
        assert out.dtype == torch.float32
        assert (out - ref).abs().max() <= 1e-3 * max(1.0, float(ref.abs().max()))
    else:
        close_bf16(out, ref, "gemm")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_focus_gather(dtype, cft):
    img = torch.rand(2, 3, 32, 48, generator=torch.Generator().manual_seed(1)).to(DEV).to(dtype)
    y = cft.ops.focus_gather(img)
    y64 = cft.ops.focus_gather(img, layout=1)
    torch.cuda.synchronize()
    x = img.float()
    ref = torch.cat([x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]], 1)   # common.py:179
    ref = ref.to(torch.bfloat16).float()
    assert y.shape == (2, 16, 16, 24)
    assert torch.equal(y[:, :12].float(), ref)
    assert (y[:, 12:] == 0).all()
    # layout 1: channel kx*16 + s = space-to-depth channel s of pixel x + kx - 1 (zero outside the image)
    assert y64.shape == (2, 64, 16, 24)
    padded = F.pad(ref, (1, 1))
    for kx in range(3):
        assert torch.equal(y64[:, kx * 16:kx * 16 + 12].float(), padded[..., kx:kx + 24]), kx
        assert (y64[:, kx * 16 + 12:kx * 16 + 16] == 0).all()
    assert (y64[:, 48:] == 0).all()


@pytest.mark.parametrize("B,Cin,Cout,H,W", [(2, 64, 64, 32, 48), (1, 64, 32, 20, 24), (2, 128, 128, 16, 16)])
def test_conv_3x1_filter(B, Cin, Cout, H, W, cft):
    """k = 3, kw = 1 (the Focus conv after the x-direction im2col): with and without the row-reuse tile shape."""
    ops = cft.ops
    x = nhwc(rnd(B, Cin, H, W, seed=1))
    w = rnd(Cout, Cin, 3, 1, seed=2, scale=1.0 / math.sqrt(Cin * 3))
    b = rnd(Cout, seed=3, scale=0.5)
    wp, bp = ops.pack_conv_weight(w, b, None, device=DEV)
    assert wp.shape == (Cout, 3, Cin)
    y = ops.conv2d(x, wp, bp, 3, 1, 1, cout=Cout, kw=1)
    y_ref_kernel = ops.conv2d(x, wp, bp, 3, 1, 1, cout=Cout, kw=1, impl="ref")
    torch.cuda.synchronize()
    ref = F.silu(F.conv2d(x.float(), wp.float().reshape(Cout, 3, 1, Cin).permute(0, 3, 1, 2), bp, padding=(1, 0)))
    close_bf16(y_ref_kernel, ref, "cuda-core ref 3x1")
    close_bf16(y, ref, "tcgen05 3x1")


@pytest.mark.parametrize("wide", [False, True])
def test_focus_module_matches_torch(wide, cft):
    """Focus = gather + tcgen05 conv == SiLU(BN(conv3x3(space_to_depth(x)))) (common.py:168-180), both layouts."""
    torch.manual_seed(0)
    m = cft.Focus(3, 64, 3).eval()
    m.wide = wide
    m.conv.bn.eps = 1e-3
    with torch.no_grad():
        m.conv.bn.running_mean.normal_(0, .1); m.conv.bn.running_var.uniform_(.5, 1.5)
        m.conv.bn.weight.uniform_(.5, 1.5); m.conv.bn.bias.normal_(0, .1)
    m = m.to(DEV)
    img = torch.rand(2, 3, 64, 96, generator=torch.Generator().manual_seed(1)).to(DEV)
    with torch.no_grad():
        y = m(img)
        x = img.to(torch.bfloat16).float()
        s2d = torch.cat([x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]], 1)
        ref = F.silu(F.batch_norm(F.conv2d(s2d, m.conv.conv.weight.to(torch.bfloat16).float(), None, padding=1),
                                  m.conv.bn.running_mean, m.conv.bn.running_var, m.conv.bn.weight, m.conv.bn.bias, False, 0., 1e-3))
    torch.cuda.synchronize()
    close_bf16(y, ref, "Focus module")


def _focus_reference(img_u8, conv_w, bn, bias=None):
    """reference Focus on the loader's bytes: x = img / 255 (test.py:107-108), space-to-depth, conv 3x3, BN, SiLU."""
    x = img_u8.float() / 255.0
    s2d = torch.cat([x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]], 1)      # common.py:179
    y = F.conv2d(s2d, conv_w, bias, padding=1)
    if bn is not None:
        y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0., bn.eps)
    return y


@pytest.mark.parametrize("B,Cout,H,W", [(2, 64, 64, 96), (1, 32, 32, 32), (2, 80, 48, 80), (3, 64, 80, 112), (1, 128, 16, 16)])
def test_focus_conv_fused_matches_reference(B, Cout, H, W, cft):
    """cft_focus_conv (uint8 image -> 6x6 stride-2 tcgen05 conv, fp16 operands) == SiLU(BN(conv3x3(space_to_depth(x/255)))).
    Tolerance: fp16 weights (2^-11 relative) on exactly represented pixels, bf16 output rounding (2^-9)."""
    torch.manual_seed(Cout + H)
    m = cft.Focus(3, Cout, 3).eval()
    m.conv.bn.eps = 1e-3
    with torch.no_grad():
        m.conv.bn.running_mean.normal_(0, .1); m.conv.bn.running_var.uniform_(.5, 1.5)
        m.conv.bn.weight.uniform_(.5, 1.5); m.conv.bn.bias.normal_(0, .1)
    m = m.to(DEV)
    # the RGB and IR halves of the loader's [B,6,H,W] tensor are consumed in place (batch stride 6*H*W)
    x6 = torch.randint(0, 256, (B, 6, H, W), dtype=torch.uint8, generator=torch.Generator().manual_seed(5)).to(DEV)
    for half in (x6[:, :3], x6[:, 3:]):
        assert cft.ops.focus_conv_supported(half, Cout, 1)
        with torch.no_grad():
            y = m(half)
            ref = F.silu(_focus_reference(half, m.conv.conv.weight, m.conv.bn))
        torch.cuda.synchronize()
        assert y.shape == (B, Cout, H // 2, W // 2)
        d = (y.float() - ref).abs()
        tol = 2e-3 + 6e-3 * ref.abs()
        assert (d <= tol).all(), f"fused Focus: max |d| {d.max().item():.4g}, worst excess {(d - tol).max().item():.4g}"


def test_focus_conv_fused_vs_two_kernel_path_and_no_act(cft):
    """Fused and gather+conv paths agree (the fused one is the more accurate: no bf16 rounding of x/255); act = none."""
    ops = cft.ops
    w = rnd(64, 12, 3, 3, seed=2, scale=1.0 / math.sqrt(108)).cpu()
    b = rnd(64, seed=3, scale=0.5).cpu()
    img = torch.randint(0, 256, (2, 3, 64, 64), dtype=torch.uint8, generator=torch.Generator().manual_seed(7)).to(DEV)
    wf, bf = ops.pack_focus_weight(w, b, None, device=DEV)
    assert wf.shape == (64, 192) and wf.dtype == torch.float16
    y = ops.focus_conv(img, wf, bf, 64, 0)
    wp, bp = ops.pack_conv_weight(w, b, None, cin_pad_to=16, device=DEV)
    y2 = ops.conv2d(ops.focus_gather(img), wp, bp, 3, 1, 0, cout=64, cin=16)
    ref = _focus_reference(img, w.to(DEV), None, b.to(DEV))
    torch.cuda.synchronize()
    e_fused = (y.float() - ref).abs().max().item()
    e_two = (y2.float() - ref).abs().max().item()
    assert e_fused <= 2e-3 + 6e-3 * ref.abs().max().item(), e_fused
    assert e_fused <= e_two + 1e-3, (e_fused, e_two)
    close_bf16(y2, ref, "two-kernel Focus")


def test_maxpool_cascade_equals_5_9_13(cft):
    ops = cft.ops
    x = nhwc(rnd(2, 64, 20, 20, seed=3))
    cat = torch.empty((2, 256, 20, 20), device=DEV, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ops.copy_into(x, cat[:, :64])
    ops.maxpool_s1(cat[:, :64], cat[:, 64:128], 5)
    ops.maxpool_s1(cat[:, 64:128], cat[:, 128:192], 5)
    ops.maxpool_s1(cat[:, 128:192], cat[:, 192:256], 5)
    direct13 = torch.empty_like(x)
    ops.maxpool_s1(x, direct13, 13)
    torch.cuda.synchronize()
    xf = x.float()
    for i, k in enumerate((5, 9, 13)):
        assert torch.equal(cat[:, 64 * (i + 1):64 * (i + 2)].float(), F.max_pool2d(xf, k, 1, k // 2)), k
    assert torch.equal(direct13.float(), F.max_pool2d(xf, 13, 1, 6))
    assert torch.equal(cat[:, :64], x)


@pytest.mark.parametrize("H,W,C", [(20, 20, 512), (32, 40, 64), (7, 5, 16)])
def test_maxpool_cascade3_is_spp(H, W, C, cft):
    """One-pass SPP pools: cascade (5,5,5) == max_pool2d k=5,9,13 (models/common.py:160-165), bit-exact."""
    ops = cft.ops
    x = nhwc(rnd(2, C, H, W, seed=4))
    cat = torch.zeros((2, 4 * C, H, W), device=DEV, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ops.copy_into(x, cat[:, :C])
    ops.maxpool_cascade3(cat[:, :C], cat, [C, 2 * C, 3 * C], [5, 5, 5])
    torch.cuda.synchronize()
    xf = x.float()
    assert torch.equal(cat[:, :C], x)
    for i, k in enumerate((5, 9, 13)):
        assert torch.equal(cat[:, C * (i + 1):C * (i + 2)].float(), F.max_pool2d(xf, k, 1, k // 2)), k


def test_upsample_add_copy(cft):
    ops = cft.ops
    a, b = nhwc(rnd(2, 64, 10, 12, seed=1)), nhwc(rnd(2, 64, 10, 12, seed=2))
    up = ops.upsample2x(a)
    s = ops.add(a, b)
    torch.cuda.synchronize()
    assert torch.equal(up.float(), F.interpolate(a.float(), scale_factor=2, mode="nearest"))
    assert torch.equal(s.float(), (a.float() + b.float()).to(torch.bfloat16).float())


CHAIN_CASES = [
    # B, C, H, W, k, s, residual, skip_out
    (2, 64, 32, 32, 3, 1, True, False),        # P2-style: row-reuse tiles, resident 3x3 weights, single CTA, N = 64
    (2, 128, 32, 32, 3, 1, True, False),       # P3-style: row-reuse tiles, CTA pairs (second GEMM as cta_group::2, M = 256)
    (1, 128, 20, 20, 3, 1, False, False),      # plain 3x3 path (20 is no multiple of 8), pairs, clipped tiles
    (2, 64, 24, 40, 1, 1, False, False),       # chained behind a flat 1x1
    (3, 128, 16, 16, 3, 1, False, True),       # y itself not written (head Bottlenecks without shortcut)
    (8, 128, 80, 80, 3, 1, True, False),       # 800 tiles: both teams, many tiles per CTA, accumulator / staging reuse
    (8, 64, 160, 160, 3, 1, True, False),      # the yolov5l P2 Bottleneck shape at batch 8
    (5, 128, 32, 24, 3, 2, False, False),      # stride-2 producer
    (5, 128, 20, 20, 3, 1, True, False),       # batch-spanning tiles (4 x 4 px x 5 images), residual, chained 1x1
]


@pytest.mark.parametrize("B,C,H,W,k,s,res,skip", CHAIN_CASES)
def test_conv_chained_1x1(B, C, H, W, k, s, res, skip, cft):
    """Back-to-back GEMM: y = SiLU(conv(x) + b) (+ residual), y2 = SiLU(W2 . bf16(y) + b2) in ONE launch (a Bottleneck's cv1
    fused into the producer of its input, models/common.py:104-106) vs fp32 torch on the same bf16-rounded operands, and vs
    the two separate launches of the same library (which round at the same places)."""
    ops = cft.ops
    x = nhwc(rnd(B, C, H, W, seed=1))
    w = rnd(C, C, k, k, seed=2, scale=1.0 / math.sqrt(C * k * k))
    b = rnd(C, seed=3, scale=0.2)
    w2 = rnd(C, C, 1, 1, seed=4, scale=1.0 / math.sqrt(C))
    b2 = rnd(C, seed=5, scale=0.2)
    ho, wo = (H + s - 1) // s, (W + s - 1) // s
    r = nhwc(rnd(B, C, ho, wo, seed=6)) if res else None
    wp, bp = ops.pack_conv_weight(w, b, None, device=DEV)
    w2p, b2p = ops.pack_conv_weight(w2, b2, None, device=DEV)
    y, y2 = ops.conv2d(x, wp, bp, k, s, 1, residual=r, cout=C, chain=(w2p, b2p, 1, None), skip_out=skip)
    y_sep = ops.conv2d(x, wp, bp, k, s, 1, residual=r, cout=C)
    y2_sep = ops.conv2d(y_sep, w2p, b2p, 1, 1, 1, cout=C)
    torch.cuda.synchronize()
    y_ref = conv_ref(x.float().cpu(), w.to(torch.bfloat16), b, k, s, 1, r.float().cpu() if res else None)
    y2_ref = conv_ref(y_ref.to(torch.bfloat16), w2.to(torch.bfloat16), b2, 1, 1, 1)
    assert (y is None) == skip
    if not skip:
        close_bf16(y.cpu(), y_ref, "chain y")
        assert torch.equal(y, y_sep)
    close_bf16(y2.cpu(), y2_ref, "chain y2")
    assert float((y2.float() - y2_sep.float()).abs().max()) <= 2e-2 * float(y2_sep.float().abs().max())


def test_c3_chain_equals_separate_launches(cft):
    """C3 with the Bottleneck cv1s chained into the 3x3 epilogues == the same C3 launching every conv separately."""
    torch.manual_seed(0)
    for c, n, shortcut, hw in ((128, 3, True, 32), (256, 4, True, 32), (256, 3, False, 16)):
        m = cft.modules.C3(c, c, n, shortcut).eval()
        with torch.no_grad():
            for mod in m.modules():
                if isinstance(mod, torch.nn.BatchNorm2d):
                    mod.running_var.uniform_(0.5, 1.5)
                    mod.running_mean.normal_(0, 0.1)
                    mod.weight.uniform_(0.5, 1.5)
                    mod.bias.normal_(0, 0.1)
        m = m.to(DEV)
        x = nhwc(rnd(2, c, hw, hw, seed=c))
        n0 = cft._lib.launch_count()
        y_chain = m(x)
        n1 = cft._lib.launch_count()
        m.chain = False
        y_sep = m(x)
        n2 = cft._lib.launch_count()
        torch.cuda.synchronize()
        assert (n2 - n1) - (n1 - n0) == n - 1                       # n - 1 launches fewer
        d = float((y_chain.float() - y_sep.float()).abs().max() / y_sep.float().abs().max())
        assert d <= 2e-2, (c, n, d)


@pytest.mark.parametrize("H,W,C", [(80, 80, 256), (20, 20, 512), (16, 20, 128), (12, 8, 64)])
def test_gpt_pool_tokens(H, W, C, cft):
    """AdaptiveAvgPool2d((8,8)) incl. the overlapping non-uniform bins of 20->8."""
    rgb, ir = nhwc(rnd(2, C, H, W, seed=1)), nhwc(rnd(2, C, H, W, seed=2))
    pos = rnd(1, 128, C, seed=3, scale=0.02).to(DEV)
    tok = cft.ops.gpt_pool_tokens(rgb, ir, pos, 8, 8)
    torch.cuda.synchronize()
    r = F.adaptive_avg_pool2d(rgb.float(), (8, 8)).view(2, C, -1)
    i = F.adaptive_avg_pool2d(ir.float(), (8, 8)).view(2, C, -1)
    ref = torch.cat([r, i], 2).permute(0, 2, 1) + pos           # common.py:615-621
    assert (tok - ref).abs().max() <= 1e-5 * max(1.0, float(ref.abs().max())) + 1e-5


@pytest.mark.parametrize("C", [128, 256, 512, 640, 1024, 1280, 2048])
def test_layernorm(C, cft):
    x = rnd(300, C, seed=1, scale=3.0).to(DEV) + 0.5
    g, b = (torch.rand(C) + 0.5).to(DEV), rnd(C, seed=2, scale=0.1).to(DEV)
    y32 = cft.ops.layernorm(x, g, b, 1e-5, out_dtype=torch.float32)
    y16 = cft.ops.layernorm(x, g, b, 1e-5)
    torch.cuda.synchronize()
    ref = F.layer_norm(x, (C,), g, b, 1e-5)
    assert (y32 - ref).abs().max() <= 2e-5 * max(1.0, float(ref.abs().max()))
    assert (y16.float() - ref).abs().max() <= 2 ** -7 * float(ref.abs().max())


@pytest.mark.parametrize("C,heads", [(128, 8), (256, 8), (512, 8), (1024, 8), (320, 8),
                                     (640, 8), (1280, 8), (384, 8), (896, 8), (1536, 8), (192, 8), (576, 8)])
def test_attention_core(C, heads, cft):
    """softmax(q k^T / sqrt(dk)) v per (image, head), 128 tokens (common.py:497-510).  Head dims 16..128 (yolov5s/l) and
    the mixed-chunk splits of the tcgen05 kernel: 80 = 64+16 and 160 = 64+64+32 (yolov5x P4/P5), 48 = 32+16,
    112 = 64+32+16, 192 = 3x64; 40 (yolov5x P3) = 32 + 16 and 24 / 72 likewise, the overhanging 8 columns zero-filled by the
    3-D TMA box (head dim, q|k|v x head, token)."""
    B, T = 3, 128
    qkv = rnd(B * T, 3 * C, seed=1).to(DEV).to(torch.bfloat16)
    out = cft.ops.attention(qkv, B, T, C, heads)
    torch.cuda.synchronize()
    dk = C // heads
    q, k, v = (qkv.float()[:, i * C:(i + 1) * C].view(B, T, heads, dk).permute(0, 2, 1, 3) for i in range(3))
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dk), -1)
    ref = (att @ v).permute(0, 2, 1, 3).reshape(B * T, C)
    assert (out.float() - ref).abs().max() <= 2e-2 * float(ref.abs().max())


@pytest.mark.parametrize("H,W", [(80, 80), (20, 20), (12, 20)])
def test_gpt_unpool_bilinear_add2_add(H, W, cft):
    """F.interpolate(mode='bilinear') (align_corners=False) + Add2 x2 + Add in one pass (common.py:626-637,229-242)."""
    B, C = 2, 64
    tok = rnd(B, 128, C, seed=1).to(DEV)
    xr, xi = nhwc(rnd(B, C, H, W, seed=2)), nhwc(rnd(B, C, H, W, seed=3))
    o_r, o_i, o_s = cft.ops.gpt_unpool(tok, H, W, 8, 8, x_rgb=xr, x_ir=xi, want_sum=True)
    u_r, u_i, _ = cft.ops.gpt_unpool(tok, H, W, 8, 8)
    torch.cuda.synchronize()
    t = tok.view(B, 2, 8, 8, C).permute(0, 1, 4, 2, 3)
    ur = F.interpolate(t[:, 0].contiguous(), size=[H, W], mode="bilinear")
    ui = F.interpolate(t[:, 1].contiguous(), size=[H, W], mode="bilinear")
    close = lambda a, r: (a.float() - r).abs().max() <= 2 ** -7 * float(r.abs().max()) + 1e-6
    assert close(u_r, ur) and close(u_i, ui)
    assert close(o_r, xr.float() + ur) and close(o_i, xi.float() + ui)
    assert close(o_s, xr.float() + ur + xi.float() + ui)


def test_detect_decode_indexing_bit_exact_and_values(cft, oracle):
    """Row order / grid / anchor assignment must be bit-exact (models/yolo_test.py:48-64)."""
    B, na, no = 2, 3, 8
    anchors = torch.tensor([[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]]).float()
    ag = anchors.view(3, 1, 3, 1, 1, 2)
    sizes = [(6, 10), (3, 5), (2, 3)]
    total = sum(na * ny * nx for ny, nx in sizes)
    for mode in ("zeros", "random"):
        z = torch.empty(B, total, no, device=DEV)
        raws, raws_cpu, row0 = [], [], 0
        for lvl, (ny, nx) in enumerate(sizes):
            head = torch.zeros(B * ny * nx, 24) if mode == "zeros" else rnd(B * ny * nx, 24, seed=lvl, scale=2.0)
            raw = cft.ops.detect_decode(head.to(DEV), B, ny, nx, na, no, [8.0, 16.0, 32.0][lvl],
                                        anchors[lvl].to(DEV), z, row0)
            raws.append(raw)
            raws_cpu.append(head.view(B, ny, nx, na, no).permute(0, 3, 1, 2, 4).contiguous())   # :48
            row0 += na * ny * nx
        torch.cuda.synchronize()
        z_ref = oracle.decode_heads(raws_cpu, ag)
        for r, rc in zip(raws, raws_cpu):
            assert torch.equal(r.cpu(), rc)                        # permute is a pure index map
        if mode == "zeros":
            assert torch.equal(z.cpu(), z_ref)                     # sigmoid(0)=.5 exact -> whole decode exact
        else:
            assert torch.allclose(z.cpu(), z_ref, rtol=1e-5, atol=1e-5)
