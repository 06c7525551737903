"""Tensor-level wrappers over the C ABI (``include/cft_b200.h``).

PyTorch is plumbing here: it owns device memory and the current stream; every arithmetic
op is a kernel of ``libcft_b200.so``.  Activations are logical NCHW tensors stored NHWC
(``torch.channels_last``) in bf16; a channel slice ``buf[:, c0:c1]`` of such a tensor is a
valid activation too (that is how producers write straight into Concat buffers).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_SILU, DT_BF16, DT_F32, ConvArgs, GptBlockArgs

__all__ = [
    "ACT_NONE", "ACT_SILU", "ACT_GELU", "empty_nhwc", "to_nhwc_bf16", "conv2d", "gemm", "focus_gather",
    "maxpool_s1", "maxpool_cascade3", "upsample2x", "add", "copy_into", "gpt_pool_tokens", "layernorm", "attention",
    "gpt_block", "gpt_block_supported", "gpt_unpool", "detect_decode", "pack_conv_weight", "pack_linear_weight",
]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise _lib.CftError(f"{what}: tensor is on {t.device}; the CFT forward path is CUDA (sm_100a) only "
                            "-- there is no CPU fallback")


def empty_nhwc(b: int, c: int, h: int, w: int, device, dtype=torch.bfloat16) -> torch.Tensor:
    return torch.empty((b, c, h, w), dtype=dtype, device=device, memory_format=torch.channels_last)


def nhwc_desc(t: torch.Tensor) -> Tuple[int, int]:
    """(data_ptr, ld) of a logical-NCHW / physical-NHWC bf16 activation (or channel slice of one)."""
    _require_cuda(t, "activation")
    b, c, h, w = t.shape
    ld = t.stride(3) if w > 1 else (t.stride(2) if h > 1 else (t.stride(0) if b > 1 else c))
    ok = (t.dtype == torch.bfloat16 and (c == 1 or t.stride(1) == 1)
          and (w == 1 or t.stride(3) == ld) and (h == 1 or t.stride(2) == w * ld)
          and (b == 1 or t.stride(0) == h * w * ld) and ld >= c)
    if not ok:
        raise _lib.CftError(f"activation must be bf16 channels_last (shape {tuple(t.shape)} strides {t.stride()} "
                            f"dtype {t.dtype})")
    return t.data_ptr(), ld


def is_nhwc_bf16(t: torch.Tensor) -> bool:
    try:
        nhwc_desc(t)
        return True
    except _lib.CftError:
        return False


def to_nhwc_bf16(t: torch.Tensor) -> torch.Tensor:
    """Module-edge conversion for tensors handed in by foreign (reference) code."""
    _require_cuda(t, "activation")
    if is_nhwc_bf16(t):
        return t
    return t.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


# ------------------------------------------------------------------------------ weights
def pack_conv_weight(w: torch.Tensor, bias: Optional[torch.Tensor], bn=None, cin_pad_to: int = 8,
                     cout_pad_to: int = 8, device=None):
    """OIHW fp32 conv weight (+ optional BatchNorm to fold, utils/torch_utils.py:181-201) ->
    (bf16 [Cout_p][k*k][Cin_p], fp32 bias [Cout_p]).  ``bn`` = (gamma, beta, mean, var, eps)."""
    w = w.detach().float()
    cout, cin, kh, kw = w.shape
    b = bias.detach().float() if bias is not None else torch.zeros(cout, device=w.device)
    if bn is not None:
        gamma, beta, mean, var, eps = bn
        scale = gamma.detach().float() / torch.sqrt(var.detach().float() + eps)
        w = w * scale.view(-1, 1, 1, 1)
        b = (b - mean.detach().float()) * scale + beta.detach().float()
    cin_p = (cin + cin_pad_to - 1) // cin_pad_to * cin_pad_to
    cout_p = (cout + cout_pad_to - 1) // cout_pad_to * cout_pad_to
    packed = torch.zeros(cout_p, kh * kw, cin_p, dtype=torch.float32, device=w.device)
    packed[:cout, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
    bias_p = torch.zeros(cout_p, dtype=torch.float32, device=w.device)
    bias_p[:cout] = b
    dev = device if device is not None else w.device
    return packed.to(device=dev, dtype=torch.bfloat16).contiguous(), bias_p.to(dev).contiguous()


def pack_linear_weight(w: torch.Tensor, bias: Optional[torch.Tensor], device=None):
    """nn.Linear [out,in] -> (bf16 [out][1][in], fp32 bias)."""
    return pack_conv_weight(w.detach().view(w.shape[0], w.shape[1], 1, 1), bias, None, device=device)


# ------------------------------------------------------------------------------ conv / gemm
def conv_chain_supported(cout: int) -> bool:
    """Can a 1x1 with ``cout`` -> ``cout`` channels ride in the epilogue of the conv that produces its input?"""
    return cout in (64, 128)


def conv2d(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], k: int, stride: int, act: int,
           out: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, cout: Optional[int] = None,
           cin: Optional[int] = None, impl: str = "tcgen05", kw: int = 0, chain=None, skip_out: bool = False):
    """y = act(conv(x, w) + bias) [+ residual] on NHWC bf16; ``out`` may be a channel-slice view.
    ``k`` is the kernel height, ``kw`` its width (0 = square).
    ``chain = (w2, bias2, act2, out2)``: additionally y2 = act2(w2 . y + bias2) (a 1x1 with Cout -> Cout channels, weights
    from ``pack_conv_weight``) computed in the same launch from the on-chip tile of y; returns ``(y, y2)``.  ``skip_out``:
    y is not written (``None`` is returned in its place)."""
    lib = _lib.lib()
    xp, ldx = nhwc_desc(x)
    b, c, h, wd = x.shape
    cin = cin if cin is not None else c
    cout = cout if cout is not None else w.shape[0]
    ho, wo = (h + stride - 1) // stride, (wd + stride - 1) // stride
    if out is None:
        out = empty_nhwc(b, cout, ho, wo, x.device)
    yp, ldy = nhwc_desc(out)
    if tuple(out.shape) != (b, cout, ho, wo):
        raise _lib.CftError(f"conv2d: out shape {tuple(out.shape)} != {(b, cout, ho, wo)}")
    a = ConvArgs()
    a.x, a.B, a.H, a.W, a.Cin, a.ldx, a.x_coff = xp, b, h, wd, cin, ldx, 0
    a.w, a.bias = w.data_ptr(), (bias.data_ptr() if bias is not None else None)
    a.Cout, a.k, a.stride, a.act = cout, k, stride, act
    if residual is not None:
        rp, ldr = nhwc_desc(residual)
        a.res, a.ldr, a.r_coff = rp, ldr, 0
    else:
        a.res, a.ldr, a.r_coff = None, 0, 0
    a.y, a.ldy, a.y_coff, a.out_dtype = yp, ldy, 0, DT_BF16
    a.kw = kw
    out2 = None
    if chain is not None:
        w2, bias2, act2, out2 = chain
        if out2 is None:
            out2 = empty_nhwc(b, cout, ho, wo, x.device)
        y2p, ldy2 = nhwc_desc(out2)
        if tuple(out2.shape) != (b, cout, ho, wo) or tuple(w2.shape[:1]) != (cout,):
            raise _lib.CftError(f"conv2d: chained 1x1 needs out2 {(b, cout, ho, wo)} and a [{cout}, 1, {cout}] weight")
        a.w2, a.bias2, a.y2, a.ldy2, a.y2_coff, a.act2 = w2.data_ptr(), (bias2.data_ptr() if bias2 is not None else None), \
            y2p, ldy2, 0, act2
        a.skip_y = 1 if skip_out else 0
    fn = lib.cft_conv2d if impl == "tcgen05" else lib.cft_conv2d_ref
    _lib.check(fn(C.byref(a), _stream()), "cft_conv2d")
    if chain is not None:
        return (None if skip_out else out), out2
    return out


def gemm(a_mat: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], act: int = ACT_NONE,
         out: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
         out_dtype: torch.dtype = torch.bfloat16, n: Optional[int] = None, impl: str = "tcgen05") -> torch.Tensor:
    """out[M,N] = act(a[M,K] @ w[N,K]^T + bias) [+ residual];  a bf16 row-major (row stride = ld),
    out/residual bf16 or fp32 (same dtype).  nn.Linear of the GPT blocks and the Detect convs."""
    lib = _lib.lib()
    _require_cuda(a_mat, "gemm A")
    m, kdim = a_mat.shape
    n = n if n is not None else w.shape[0]
    if a_mat.dtype != torch.bfloat16 or a_mat.stride(1) != 1:
        raise _lib.CftError("gemm: A must be bf16 with unit column stride")
    if out is None:
        out = torch.empty((m, n), dtype=out_dtype, device=a_mat.device)
    if out.stride(1) != 1 or out.dtype not in (torch.bfloat16, torch.float32):
        raise _lib.CftError("gemm: bad out")
    args = ConvArgs()
    args.x, args.B, args.H, args.W, args.Cin, args.ldx, args.x_coff = a_mat.data_ptr(), 1, 1, m, kdim, a_mat.stride(0), 0
    args.w, args.bias = w.data_ptr(), (bias.data_ptr() if bias is not None else None)
    args.Cout, args.k, args.stride, args.act = n, 1, 1, act
    if residual is not None:
        if residual.dtype != out.dtype or residual.stride(1) != 1:
            raise _lib.CftError("gemm: residual must match out dtype")
        args.res, args.ldr, args.r_coff = residual.data_ptr(), residual.stride(0), 0
    else:
        args.res, args.ldr, args.r_coff = None, 0, 0
    args.y, args.ldy, args.y_coff = out.data_ptr(), out.stride(0), 0
    args.out_dtype = DT_F32 if out.dtype == torch.float32 else DT_BF16
    args.kw = 0
    fn = lib.cft_conv2d if impl == "tcgen05" else lib.cft_conv2d_ref
    _lib.check(fn(C.byref(args), _stream()), "cft_conv2d(gemm)")
    return out


# ------------------------------------------------------------------------------ movers
def focus_gather(img: torch.Tensor, layout: int = 0) -> torch.Tensor:
    """NCHW image [B,3,H,W] -> NHWC bf16 at half resolution: layout 0 = [B,16,H/2,W/2] (12 used), layout 1 =
    [B,64,H/2,W/2] x-direction im2col (patches of x-1, x, x+1 side by side; 48 used).  fp32 / bf16 values in [0,1],
    or uint8 (the loader's wire format, scaled by 1/255 in the kernel).  The batch stride may be larger than 3*H*W
    (the RGB / IR halves of the loader's [B,6,H,W] tensor) -- no copy is made for such views."""
    lib = _lib.lib()
    _require_cuda(img, "image")
    if img.dtype not in (torch.float32, torch.bfloat16, torch.uint8):
        img = img.float()
    b, c, h, w = img.shape
    if c != 3:
        raise _lib.CftError(f"focus_gather expects 3 input channels, got {c}")
    if not (img.stride(3) == 1 and img.stride(2) == w and img.stride(1) == h * w and img.stride(0) % 2 == 0):
        img = img.contiguous()
    bstride = img.stride(0) if b > 1 else 3 * h * w
    dt = {torch.float32: DT_F32, torch.bfloat16: DT_BF16, torch.uint8: _lib.DT_U8}[img.dtype]
    y = empty_nhwc(b, 64 if layout == 1 else 16, h // 2, w // 2, img.device)
    _lib.check(lib.cft_focus_gather(img.data_ptr(), dt, b, h, w, bstride, layout, y.data_ptr(), _stream()),
               "cft_focus_gather")
    return y


def pack_focus_weight(w: torch.Tensor, bias: Optional[torch.Tensor], bn=None, device=None):
    """Focus conv weight [Cout, 12, 3, 3] (+ BatchNorm to fold) -> (fp16 [Cout_p][192], fp32 bias [Cout_p]) for
    ``focus_conv``: the 3x3 filter over the 12 space-to-depth channels (channel (gy + 2 gx) * 3 + c,
    models/common.py:179) re-indexed as a 6x6 stride-2 filter on the image, K = ((c * 6 + r) * 8 + q) with r = 2 ky + gy,
    q = 2 kx + gx; q = 6, 7 and K >= 144 are zero padding.  Cout is padded to a multiple of 16."""
    w = w.detach().float()
    cout, cin, kh, kw = w.shape
    if cin != 12 or kh != 3 or kw != 3:
        raise _lib.CftError(f"pack_focus_weight expects a [Cout,12,3,3] weight, got {tuple(w.shape)}")
    b = bias.detach().float() if bias is not None else torch.zeros(cout, device=w.device)
    if bn is not None:
        gamma, beta, mean, var, eps = bn
        scale = gamma.detach().float() / torch.sqrt(var.detach().float() + eps)
        w = w * scale.view(-1, 1, 1, 1)
        b = (b - mean.detach().float()) * scale + beta.detach().float()
    cout_p = (cout + 15) // 16 * 16
    w6 = w.view(cout, 2, 2, 3, 3, 3)                       # [o, gx, gy, c, ky, kx]  (group g = gy + 2 gx)
    w6 = w6.permute(0, 3, 4, 2, 5, 1)                      # [o, c, ky, gy, kx, gx]
    w6 = w6.reshape(cout, 3, 6, 6)                         # [o, c, r = 2 ky + gy, q = 2 kx + gx]
    packed = torch.zeros(cout_p, 24, 8, dtype=torch.float32, device=w.device)
    packed[:cout, :18, :6] = w6.reshape(cout, 18, 6)
    bias_p = torch.zeros(cout_p, dtype=torch.float32, device=w.device)
    bias_p[:cout] = b
    dev = device if device is not None else w.device
    return packed.view(cout_p, 192).to(device=dev, dtype=torch.float16).contiguous(), bias_p.to(dev).contiguous()


def focus_conv_supported(img: torch.Tensor, cout: int, act: int) -> bool:
    """The fused kernel reads the loader's uint8 image through TMA: 16-byte aligned rows and image planes."""
    if img.dtype != torch.uint8 or img.dim() != 4 or img.shape[1] != 3:
        return False
    b, _, h, w = img.shape
    ok_layout = img.stride(3) == 1 and img.stride(2) == w and img.stride(1) == h * w and (b == 1 or img.stride(0) % 16 == 0)
    return (ok_layout and h % 2 == 0 and w % 16 == 0 and (h * w) % 16 == 0 and img.data_ptr() % 16 == 0 and cout <= 128
            and act in (_lib.ACT_NONE, _lib.ACT_SILU))


def focus_conv(img: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, cout: int, act: int,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused Focus (space-to-depth + 3x3 conv + bias + SiLU) from the uint8 image [B,3,H,W] (a view of the loader's
    [B,6,H,W] tensor is fine) to NHWC bf16 [B,cout,H/2,W/2]; ``w, bias`` from ``pack_focus_weight``."""
    lib = _lib.lib()
    _require_cuda(img, "image")
    if not focus_conv_supported(img, cout, act):
        raise _lib.CftError("focus_conv: unsupported image layout / width / activation (use focus_gather + conv2d)")
    b, _, h, wd = img.shape
    if out is None:
        out = empty_nhwc(b, cout, h // 2, wd // 2, img.device)
    yp, ldy = nhwc_desc(out)
    bstride = img.stride(0) if b > 1 else 3 * h * wd
    _lib.check(lib.cft_focus_conv(img.data_ptr(), b, h, wd, bstride, w.data_ptr(), bias.data_ptr(), w.shape[0], act,
                                  yp, ldy, 0, _stream()), "cft_focus_conv")
    return out


def maxpool_s1(x: torch.Tensor, out: torch.Tensor, k: int) -> torch.Tensor:
    lib = _lib.lib()
    xp, ldx = nhwc_desc(x)
    yp, ldy = nhwc_desc(out)
    b, c, h, w = x.shape
    _lib.check(lib.cft_maxpool_s1(xp, ldx, 0, yp, ldy, 0, b, h, w, c, k, _stream()), "cft_maxpool_s1")
    return out


def maxpool_cascade3(x: torch.Tensor, cat: torch.Tensor, coffs, ks) -> torch.Tensor:
    """Three stride-1 max pools in cascade (windows ks) of ``x``; stage i is written to channels
    [coffs[i], coffs[i]+C) of ``cat`` (x may itself be a slice of ``cat``)."""
    lib = _lib.lib()
    xp, ldx = nhwc_desc(x)
    yp, ldy = nhwc_desc(cat)
    b, c, h, w = x.shape
    _lib.check(lib.cft_maxpool_cascade3(xp, ldx, 0, yp, ldy, coffs[0], coffs[1], coffs[2], b, h, w, c,
                                        ks[0], ks[1], ks[2], _stream()), "cft_maxpool_cascade3")
    return cat


def upsample2x(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.lib()
    xp, ldx = nhwc_desc(x)
    b, c, h, w = x.shape
    if out is None:
        out = empty_nhwc(b, c, 2 * h, 2 * w, x.device)
    yp, ldy = nhwc_desc(out)
    _lib.check(lib.cft_upsample2x(xp, ldx, 0, yp, ldy, 0, b, h, w, c, _stream()), "cft_upsample2x")
    return out


def add(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.lib()
    ap, lda = nhwc_desc(a)
    bp, ldb = nhwc_desc(b)
    n, c, h, w = a.shape
    if out is None:
        out = empty_nhwc(n, c, h, w, a.device)
    yp, ldy = nhwc_desc(out)
    _lib.check(lib.cft_add(ap, lda, 0, bp, ldb, 0, yp, ldy, 0, n * h * w, c, _stream()), "cft_add")
    return out


def copy_into(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    lib = _lib.lib()
    xp, ldx = nhwc_desc(x)
    yp, ldy = nhwc_desc(out)
    n, c, h, w = x.shape
    _lib.check(lib.cft_copy(xp, ldx, 0, yp, ldy, 0, n * h * w, c, _stream()), "cft_copy")
    return out


# ------------------------------------------------------------------------------ GPT glue
def gpt_pool_tokens(rgb: torch.Tensor, ir: torch.Tensor, pos_emb: torch.Tensor, va: int, ha: int) -> torch.Tensor:
    lib = _lib.lib()
    rp, ldr = nhwc_desc(rgb)
    ip, ldi = nhwc_desc(ir)
    b, c, h, w = rgb.shape
    tok = torch.empty((b, 2 * va * ha, c), dtype=torch.float32, device=rgb.device)
    _lib.check(lib.cft_gpt_pool_tokens(rp, ldr, 0, ip, ldi, 0, b, h, w, c, va, ha, pos_emb.data_ptr(),
                                       tok.data_ptr(), _stream()), "cft_gpt_pool_tokens")
    return tok


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
              out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    lib = _lib.lib()
    _require_cuda(x, "layernorm input")
    c = x.shape[-1]
    rows = x.numel() // c
    y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    _lib.check(lib.cft_layernorm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), float(eps), rows, c, y.data_ptr(),
                                 DT_F32 if out_dtype == torch.float32 else DT_BF16, _stream()), "cft_layernorm")
    return y


def attention(qkv: torch.Tensor, b: int, t: int, c: int, heads: int) -> torch.Tensor:
    lib = _lib.lib()
    out = torch.empty((b * t, c), dtype=torch.bfloat16, device=qkv.device)
    _lib.check(lib.cft_attention(qkv.data_ptr(), out.data_ptr(), b, t, c, heads, _stream()), "cft_attention")
    return out


def gpt_block_supported(b: int, d: int, heads: int, tokens: int) -> bool:
    """True when the fused transformer-stack kernel (``cft_gpt_block``) covers this shape."""
    return bool(_lib.lib().cft_gpt_block_supported(b, d, heads, tokens))


def gpt_block(tok: torch.Tensor, w: dict, heads: int, cluster: int = 0, debug_x: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ln_f(trans_blocks(tok)) in ONE launch (models/common.py:622,625).  ``tok`` fp32 [B,128,d]; ``w``: the stacked
    per-layer parameters built by ``modules.GPT._weights`` (key ``"stack"``)."""
    lib = _lib.lib()
    _require_cuda(tok, "gpt_block tokens")
    b, t, d = tok.shape
    if tok.dtype != torch.float32 or not tok.is_contiguous():
        raise _lib.CftError("gpt_block: tokens must be contiguous fp32 [B,128,d]")
    if out is None:
        out = torch.empty_like(tok)
    need = int(lib.cft_gpt_block_workspace_bytes(b, d))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=tok.device)
    a = GptBlockArgs()
    a.B, a.tokens, a.d, a.heads, a.layers, a.cluster = b, t, d, heads, w["layers"], cluster
    for k in ("wqkv", "bqkv", "wo", "bo", "w1", "b1", "w2", "b2", "ln1_g", "ln1_b", "ln2_g", "ln2_b", "lnf_g", "lnf_b"):
        setattr(a, k, w[k].data_ptr())
    a.eps1, a.eps2, a.epsf = w["eps1"], w["eps2"], w["epsf"]
    a.x_in, a.x_out = tok.data_ptr(), out.data_ptr()
    a.workspace, a.workspace_bytes = workspace.data_ptr(), workspace.numel()
    a.debug_x = debug_x.data_ptr() if debug_x is not None else None
    _lib.check(lib.cft_gpt_block(C.byref(a), _stream()), "cft_gpt_block")
    out._cft_ws = workspace          # the launch is asynchronous: keep the scratch alive with its result
    return out


def gpt_unpool(tok: torch.Tensor, h: int, w: int, va: int, ha: int, x_rgb: Optional[torch.Tensor] = None,
               x_ir: Optional[torch.Tensor] = None, want_sum: bool = False,
               out_rgb: Optional[torch.Tensor] = None, out_ir: Optional[torch.Tensor] = None,
               out_sum: Optional[torch.Tensor] = None):
    lib = _lib.lib()
    b, _, c = tok.shape
    dev = tok.device
    out_rgb = out_rgb if out_rgb is not None else empty_nhwc(b, c, h, w, dev)
    out_ir = out_ir if out_ir is not None else empty_nhwc(b, c, h, w, dev)
    if want_sum and out_sum is None:
        out_sum = empty_nhwc(b, c, h, w, dev)
    xr = nhwc_desc(x_rgb) if x_rgb is not None else (None, 0)
    xi = nhwc_desc(x_ir) if x_ir is not None else (None, 0)
    orp, ori = nhwc_desc(out_rgb), nhwc_desc(out_ir)
    osum = nhwc_desc(out_sum) if out_sum is not None else (None, 0)
    _lib.check(lib.cft_gpt_unpool(tok.data_ptr(), b, h, w, c, va, ha, xr[0], xr[1], 0, xi[0], xi[1], 0,
                                  orp[0], orp[1], 0, ori[0], ori[1], 0, osum[0], osum[1], 0, _stream()),
               "cft_gpt_unpool")
    return out_rgb, out_ir, out_sum


def detect_decode(head: torch.Tensor, b: int, ny: int, nx: int, na: int, no: int, stride: float,
                  anchors_px: torch.Tensor, z: torch.Tensor, z_row0: int) -> torch.Tensor:
    lib = _lib.lib()
    raw = torch.empty((b, na, ny, nx, no), dtype=torch.float32, device=head.device)
    _lib.check(lib.cft_detect_decode(head.data_ptr(), head.stride(0), b, ny, nx, na, no, float(stride),
                                     anchors_px.data_ptr(), raw.data_ptr(), z.data_ptr(), z.shape[1], z_row0,
                                     _stream()), "cft_detect_decode")
    return raw
