"""Drop-in ``nn.Module`` replacements for the hot-path classes of the reference.

Same class names, constructor signatures, sub-module/parameter names (hence identical
``state_dict`` keys) and call protocol as reference ``models/common.py`` and
``models/yolo_test.py``, so that the reference's ``parse_model`` (``eval`` of the yaml module
names, models/yolo_test.py:488) can instantiate them unchanged -- see ``model.install``.
The arithmetic is NOT the reference's PyTorch ops: every ``forward`` drives kernels of
``libcft_b200.so`` (tcgen05 implicit-GEMM convs, fused movers) through ``ops``.  Forward only
(eval semantics: BatchNorm running statistics are folded into the conv weights, dropout = identity).

Activations between modules: logical NCHW, physical NHWC (channels_last) bf16.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops
from ._lib import CftError

ACT_NONE, ACT_SILU, ACT_GELU = ops.ACT_NONE, ops.ACT_SILU, ops.ACT_GELU


def autopad(k, p=None):  # reference models/common.py:24-28
    if p is None:
        p = k // 2 if isinstance(k, int) else [x // 2 for x in k]
    return p


def _versions(*tensors):
    return tuple((t.data_ptr(), t._version, t.device) for t in tensors if t is not None)


class _Packed:
    """Cache of kernel-layout weights, rebuilt when a source parameter changes."""

    def __init__(self):
        self.key = None
        self.value = None

    def get(self, key, builder):
        if self.key != key:
            self.value = builder()
            self.key = key
        return self.value


def _require_eval_bn(bn):
    """The kernels fold BatchNorm's RUNNING statistics into the conv weights (eval semantics).  A BatchNorm in training
    mode would need batch statistics (models/common.py:42 under train.py): not built -- fail loudly instead of silently
    computing the eval result."""
    if bn is not None and bn.training:
        raise CftError("BatchNorm in training mode (batch statistics) is not implemented on the B200 forward path: "
                       "call model.eval() -- the kernels fold the running statistics (utils/torch_utils.py:181-201)")


class Conv(nn.Module):
    """reference models/common.py:36-50: SiLU(BN(conv2d(x))) -- one fused tcgen05 kernel launch."""

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p), groups=g, bias=False)
        self.bn = nn.BatchNorm2d(c2)
        self.act = nn.SiLU() if act is True else (act if isinstance(act, nn.Module) else nn.Identity())
        self._packed = _Packed()

    # -- kernel-side view of the parameters
    def _check(self):
        c = self.conv
        k, s = c.kernel_size[0], c.stride[0]
        if c.groups != 1 or c.kernel_size[0] != c.kernel_size[1] or k not in (1, 3) or s not in (1, 2) \
                or c.padding[0] != k // 2 or (s == 2 and k != 3):
            raise CftError(f"Conv config outside the CFTx3 hot path: {c}")
        if isinstance(self.act, nn.SiLU):
            return k, s, ACT_SILU
        if isinstance(self.act, nn.Identity):
            return k, s, ACT_NONE
        raise CftError(f"unsupported activation {self.act}")

    def folded(self, device):
        """(packed bf16 weight, fp32 bias) with BN folded (utils/torch_utils.py:181-201)."""
        bn = getattr(self, "bn", None)
        _require_eval_bn(bn)
        srcs = [self.conv.weight, self.conv.bias] + ([bn.weight, bn.bias, bn.running_mean, bn.running_var] if bn is not None else [])

        def build():
            bnp = (bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps) if bn is not None else None
            return ops.pack_conv_weight(self.conv.weight, self.conv.bias, bnp, device=device)
        return self._packed.get((_versions(*srcs), str(device)), build)

    def forward(self, x, out=None, residual=None):
        k, s, act = self._check()
        x = ops.to_nhwc_bf16(x)
        w, b = self.folded(x.device)
        return ops.conv2d(x, w, b, k, s, act, out=out, residual=residual, cout=self.conv.out_channels)

    def fuseforward(self, x, out=None, residual=None):  # reference models/common.py:49-50
        # the reference's Model.fuse() rebinds `m.forward = m.fuseforward` (models/yolo_test.py:302): go through the class,
        # not through self.forward, and keep the keyword arguments the sibling modules pass (out= / residual=)
        return Conv.forward(self, x, out=out, residual=residual)


class Focus(nn.Module):
    """reference models/common.py:168-180.  uint8 images (the loader's wire format) take the FUSED kernel: space-to-depth
    + 3x3 conv + BN + SiLU as one 6x6 stride-2 tcgen05 conv on the raw image (``cft_focus_conv``).  fp32 / bf16 images
    (already scaled to [0,1]) take the two-kernel path: space-to-depth gather (16 channels, 12 used) + 3x3 conv in
    row-reuse mode; ``wide = True`` selects the x-direction-im2col layout (64 channels) + 3x1 conv for that path."""
    wide = False
    fused = os.environ.get("CFT_NO_FUSED_FOCUS") is None

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, act=True):
        super().__init__()
        self.conv = Conv(c1 * 4, c2, k, s, p, g, act)
        self._packed = _Packed()
        self._packed_fused = _Packed()     # separate caches: a captured CUDA graph keeps pointing at the tensors of its path

    def forward(self, x, out=None):
        if x.shape[1] != 3:
            raise CftError("Focus: the two-stream path feeds 3-channel images (models/yolo_test.py:499-500)")
        ops._require_cuda(x, "Focus input")
        cv = self.conv
        k, s, act = cv._check()
        bn = getattr(cv, "bn", None)
        _require_eval_bn(bn)
        srcs = [cv.conv.weight, cv.conv.bias] + ([bn.weight, bn.bias, bn.running_mean, bn.running_var] if bn is not None else [])
        wide = bool(self.wide) and k == 3 and s == 1

        def build():
            bnp = (bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps) if bn is not None else None
            w, b = ops.pack_conv_weight(cv.conv.weight, cv.conv.bias, bnp, cin_pad_to=16, device=x.device)
            if wide:                       # [Cout][ky*3+kx][16] -> [Cout][ky][kx*16 + c] (+16 zero columns)
                co = w.shape[0]
                w3 = torch.zeros(co, 3, 64, dtype=w.dtype, device=w.device)
                w3[:, :, :48] = w.view(co, 3, 3, 16).reshape(co, 3, 48)
                w = w3.contiguous()
            return w, b
        cout = cv.conv.out_channels
        if self.fused and k == 3 and s == 1 and cout % 16 == 0 and ops.focus_conv_supported(x, cout, act):
            def build_fused():
                bnp = (bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps) if bn is not None else None
                return ops.pack_focus_weight(cv.conv.weight, cv.conv.bias, bnp, device=x.device)
            wf, bf = self._packed_fused.get((_versions(*srcs), str(x.device)), build_fused)
            return ops.focus_conv(x, wf, bf, cout, act, out=out)
        w, b = self._packed.get((_versions(*srcs), str(x.device), wide), build)
        if wide:
            g = ops.focus_gather(x, layout=1)
            return ops.conv2d(g, w, b, 3, 1, act, out=out, cout=cv.conv.out_channels, cin=64, kw=1)
        g = ops.focus_gather(x, layout=0)
        return ops.conv2d(g, w, b, k, s, act, out=out, cout=cv.conv.out_channels, cin=16)


class Bottleneck(nn.Module):
    """reference models/common.py:99-109: x + cv2(cv1(x)); the residual add rides in cv2's epilogue."""

    def __init__(self, c1, c2, shortcut=True, g=1, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_, c2, 3, 1, g=g)
        self.add = shortcut and c1 == c2

    def forward(self, x, out=None):
        x = ops.to_nhwc_bf16(x)
        return self.cv2(self.cv1(x), out=out, residual=x if self.add else None)


class C3(nn.Module):
    """reference models/common.py:131-143: cv3(cat(m(cv1(x)), cv2(x))).

    cv1 and cv2 (same input, both 1x1+SiLU) run as ONE GEMM with N = 2c_ whose output is already
    the concat buffer; the last Bottleneck writes m(...) over the cv1 half, so no cat is materialised.
    """

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*[Bottleneck(c_, c_, shortcut, g, e=1.0) for _ in range(n)])
        self._packed = _Packed()

    def _cv12(self, device):
        srcs = []
        for cv in (self.cv1, self.cv2):
            bn = getattr(cv, "bn", None)
            srcs += [cv.conv.weight, cv.conv.bias] + ([bn.weight, bn.bias, bn.running_mean, bn.running_var] if bn is not None else [])

        def build():
            w1, b1 = self.cv1.folded(device)
            w2, b2 = self.cv2.folded(device)
            return torch.cat([w1, w2], 0).contiguous(), torch.cat([b1, b2], 0).contiguous()
        return self._packed.get((_versions(*srcs), str(device)), build)

    def forward(self, x, out=None):
        x = ops.to_nhwc_bf16(x)
        self.cv1._check(), self.cv2._check()
        _require_eval_bn(getattr(self.cv1, "bn", None)), _require_eval_bn(getattr(self.cv2, "bn", None))   # also when the
        c_ = self.cv1.conv.out_channels                                    # packed cv1|cv2 weights are already cached
        w12, b12 = self._cv12(x.device)
        cat = ops.conv2d(x, w12, b12, 1, 1, ACT_SILU, cout=2 * c_)        # [cv1(x) | cv2(x)]
        a = cat[:, :c_]
        n = len(self.m)
        if self.chain and n >= 2 and ops.conv_chain_supported(c_) and all(self._chainable(b, c_) for b in self.m):
            # every Bottleneck's cv1 (1x1) except the first rides in the epilogue of the previous Bottleneck's cv2 (3x3) as a
            # back-to-back GEMM on the on-chip output tile: n - 1 launches and as many re-reads of `a` fewer per C3
            t = self.m[0].cv1(a)
            for j, blk in enumerate(self.m):
                w3, b3 = blk.cv2.folded(x.device)
                res = a if blk.add else None
                if j == n - 1:
                    a = ops.conv2d(t, w3, b3, 3, 1, ACT_SILU, out=cat[:, :c_], residual=res, cout=c_)
                else:
                    nxt = self.m[j + 1]
                    w1, b1 = nxt.cv1.folded(x.device)
                    a, t = ops.conv2d(t, w3, b3, 3, 1, ACT_SILU, residual=res, cout=c_, chain=(w1, b1, ACT_SILU, None),
                                      skip_out=not nxt.add)         # a' is only read again as the next residual
            return self.cv3(cat, out=out)
        for j, blk in enumerate(self.m):
            a = blk(a, out=cat[:, :c_] if j == n - 1 else None)
        return self.cv3(cat, out=out)

    # Bottleneck cv1 fused into the preceding 3x3 (csrc/conv_tcgen05.cu, chain mode); CFT_NO_CONV_CHAIN=1 / chain = False
    # launches every 1x1 separately
    chain = os.environ.get("CFT_NO_CONV_CHAIN") is None

    @staticmethod
    def _chainable(blk, c_):
        try:
            k1, s1, a1 = blk.cv1._check()
            k2, s2, a2 = blk.cv2._check()
        except CftError:
            return False
        _require_eval_bn(getattr(blk.cv1, "bn", None)), _require_eval_bn(getattr(blk.cv2, "bn", None))
        return (k1, s1, a1, k2, s2, a2) == (1, 1, ACT_SILU, 3, 1, ACT_SILU) and blk.cv1.conv.in_channels == c_ \
            and blk.cv1.conv.out_channels == c_ and blk.cv2.conv.out_channels == c_


class SPP(nn.Module):
    """reference models/common.py:154-165. cv1 writes into the first quarter of the concat buffer;
    the 5/9/13 pools are a 5x5 cascade (max-pool stride 1 with -inf padding composes exactly)."""

    def __init__(self, c1, c2, k=(5, 9, 13)):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * (len(k) + 1), c2, 1, 1)
        self.m = nn.ModuleList([nn.MaxPool2d(kernel_size=x, stride=1, padding=x // 2) for x in k])

    def forward(self, x, out=None):
        x = ops.to_nhwc_bf16(x)
        b, _, h, w = x.shape
        c_ = self.cv1.conv.out_channels
        ks = [m.kernel_size if isinstance(m.kernel_size, int) else m.kernel_size[0] for m in self.m]
        cat = ops.empty_nhwc(b, c_ * (len(ks) + 1), h, w, x.device)
        self.cv1(x, out=cat[:, :c_])
        steps = [ks[0]] + [ks[i] - ks[i - 1] + 1 for i in range(1, len(ks))]    # pool_k = pool_step(pool_prev)
        if len(ks) == 3 and all(st >= 1 and st % 2 == 1 for st in steps) and c_ % 16 == 0 and h * w * 64 <= 200 * 1024:
            ops.maxpool_cascade3(cat[:, :c_], cat, [c_, 2 * c_, 3 * c_], steps)       # one smem-staged pass
        else:
            prev_k, prev = 1, cat[:, :c_]
            for i, k in enumerate(ks):
                dst = cat[:, (i + 1) * c_:(i + 2) * c_]
                step = k - prev_k + 1
                if i > 0 and step >= 1 and step % 2 == 1 and k > prev_k:
                    ops.maxpool_s1(prev, dst, step)
                else:
                    ops.maxpool_s1(cat[:, :c_], dst, k)
                prev_k, prev = k, dst
        return self.cv2(cat, out=out)


class Concat(nn.Module):
    """reference models/common.py:211-219. Free when the inputs are already adjacent channel slices of
    one buffer (the model's planner arranges that); otherwise two slice copies."""

    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    def forward(self, x: Sequence[torch.Tensor], out=None):
        if self.d != 1:
            raise CftError("Concat: only the channel dimension is on the hot path")
        xs = [ops.to_nhwc_bf16(t) for t in x]
        fused = _adjacent_slices(xs)
        if fused is not None and out is None:
            return fused
        b, _, h, w = xs[0].shape
        ctot = sum(t.shape[1] for t in xs)
        if out is None:
            out = ops.empty_nhwc(b, ctot, h, w, xs[0].device)
        c0 = 0
        for t in xs:
            ops.copy_into(t, out[:, c0:c0 + t.shape[1]])
            c0 += t.shape[1]
        return out


def _adjacent_slices(xs: List[torch.Tensor]) -> Optional[torch.Tensor]:
    """If xs are consecutive channel slices covering one NHWC buffer, return that buffer as a tensor."""
    base = getattr(xs[0], "_cft_base", None)
    if base is None:
        return None
    off = 0
    for t in xs:
        if getattr(t, "_cft_base", None) is not base or getattr(t, "_cft_coff", -1) != off:
            return None
        off += t.shape[1]
    return base if off == base.shape[1] else None


def concat_slot(base: torch.Tensor, c0: int, c1: int) -> torch.Tensor:
    """Channel-slice view of a concat buffer, tagged so Concat can recognise it."""
    v = base[:, c0:c1]
    v._cft_base, v._cft_coff = base, c0
    return v


class Add(nn.Module):
    """reference models/common.py:222-229"""

    def __init__(self, arg):
        super().__init__()
        self.arg = arg

    def forward(self, x, out=None):
        return ops.add(ops.to_nhwc_bf16(x[0]), ops.to_nhwc_bf16(x[1]), out=out)


class Add2(nn.Module):
    """reference models/common.py:232-243"""

    def __init__(self, c1, index):
        super().__init__()
        self.index = index

    def forward(self, x, out=None):
        if self.index not in (0, 1):
            raise CftError("Add2: index must be 0 or 1")
        return ops.add(ops.to_nhwc_bf16(x[0]), ops.to_nhwc_bf16(x[1][self.index]), out=out)


class Upsample(nn.Upsample):
    """nn.Upsample(None, 2, 'nearest') of the yaml head (rows 33/37) on the NHWC mover kernel."""

    def forward(self, x, out=None):
        if self.mode != "nearest" or float(self.scale_factor) != 2.0 or self.size is not None:
            raise CftError("Upsample: only nearest x2 is on the hot path")
        return ops.upsample2x(ops.to_nhwc_bf16(x), out=out)


# ------------------------------------------------------------------------------ CFT / GPT
class SelfAttention(nn.Module):
    """reference models/common.py:430-513 (parameter container; math runs in GPT.forward)."""

    def __init__(self, d_model, d_k, d_v, h, attn_pdrop=.1, resid_pdrop=.1):
        super().__init__()
        assert d_k % h == 0
        self.d_model, self.d_k, self.d_v, self.h = d_model, d_model // h, d_model // h, h
        self.que_proj = nn.Linear(d_model, h * self.d_k)
        self.key_proj = nn.Linear(d_model, h * self.d_k)
        self.val_proj = nn.Linear(d_model, h * self.d_v)
        self.out_proj = nn.Linear(h * self.d_v, d_model)
        self.attn_drop = nn.Dropout(attn_pdrop)
        self.resid_drop = nn.Dropout(resid_pdrop)


class myTransformerBlock(nn.Module):
    """reference models/common.py:516-546 (parameter container)."""

    def __init__(self, d_model, d_k, d_v, h, block_exp, attn_pdrop, resid_pdrop):
        super().__init__()
        self.ln_input = nn.LayerNorm(d_model)
        self.ln_output = nn.LayerNorm(d_model)
        self.sa = SelfAttention(d_model, d_k, d_v, h, attn_pdrop, resid_pdrop)
        self.mlp = nn.Sequential(
            nn.Linear(d_model, block_exp * d_model),
            nn.GELU(),
            nn.Linear(block_exp * d_model, d_model),
            nn.Dropout(resid_pdrop),
        )


class GPT(nn.Module):
    """reference models/common.py:549-639: the Cross-Modality Fusion Transformer block.

    tokeniser kernel (8x8 adaptive avg-pool of both modalities + pos_emb, fp32 residual stream)
    -> 8 x [LN -> QKV GEMM (N=3d) -> 128-token attention -> out-proj GEMM (+residual)
            -> LN -> MLP-up GEMM (+GELU) -> MLP-down GEMM (+residual)]
    -> ln_f -> bilinear un-pool kernel.  ``forward_fused`` additionally folds the two Add2 and the Add
    that follow every GPT in the x3 graphs into the un-pool pass.
    """

    def __init__(self, d_model, h=8, block_exp=4, n_layer=8, vert_anchors=8, horz_anchors=8,
                 embd_pdrop=0.1, attn_pdrop=0.1, resid_pdrop=0.1):
        super().__init__()
        self.n_embd = d_model
        self.vert_anchors, self.horz_anchors = vert_anchors, horz_anchors
        self.h = h
        self.pos_emb = nn.Parameter(torch.zeros(1, 2 * vert_anchors * horz_anchors, self.n_embd))
        self.trans_blocks = nn.Sequential(*[myTransformerBlock(d_model, d_model, d_model, h, block_exp, attn_pdrop,
                                                               resid_pdrop) for _ in range(n_layer)])
        self.ln_f = nn.LayerNorm(self.n_embd)
        self.drop = nn.Dropout(embd_pdrop)
        self.avgpool = nn.AdaptiveAvgPool2d((self.vert_anchors, self.horz_anchors))
        self.apply(self._init_weights)
        self._packed = _Packed()

    @staticmethod
    def _init_weights(module):  # reference models/common.py:583-591
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=0.02)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)

    # one-launch transformer stack (csrc/cft_block.cu); CFT_NO_FUSED_BLOCK=1 / fused_block = False -> 7 launches per layer.
    # Measured on B200 (profiles/r02_block_*.txt): the cluster-per-image kernel is a per-CTA latency chain -- 1.56x faster
    # than the per-op path at d = 256 (any batch), equal at d = 512 / batch 32 and slower at d = 512 / small batch, so it
    # is used up to fused_block_max_d; wider blocks keep the batched GEMM launches.
    fused_block = os.environ.get("CFT_NO_FUSED_BLOCK") is None
    fused_block_max_d = int(os.environ.get("CFT_FUSED_BLOCK_MAX_D", "256"))

    def _weights(self, device):
        srcs = [p for p in self.parameters()]

        def build():
            layers = []
            f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
            for blk in self.trans_blocks:
                sa = blk.sa
                wqkv = torch.cat([sa.que_proj.weight, sa.key_proj.weight, sa.val_proj.weight], 0)
                bqkv = torch.cat([sa.que_proj.bias, sa.key_proj.bias, sa.val_proj.bias], 0)
                layers.append({
                    "ln1": (f32(blk.ln_input.weight), f32(blk.ln_input.bias), blk.ln_input.eps),
                    "ln2": (f32(blk.ln_output.weight), f32(blk.ln_output.bias), blk.ln_output.eps),
                    "qkv": ops.pack_linear_weight(wqkv, bqkv, device=device),
                    "out": ops.pack_linear_weight(sa.out_proj.weight, sa.out_proj.bias, device=device),
                    "up": ops.pack_linear_weight(blk.mlp[0].weight, blk.mlp[0].bias, device=device),
                    "down": ops.pack_linear_weight(blk.mlp[2].weight, blk.mlp[2].bias, device=device),
                })
            # the same parameters stacked over the layers: one tensor per kind for the one-launch kernel (cft_gpt_block)
            cat = lambda ts: torch.cat(list(ts), 0).contiguous()
            stack = {
                "layers": len(layers),
                "wqkv": cat(L["qkv"][0].view(L["qkv"][0].shape[0], -1) for L in layers), "bqkv": cat(L["qkv"][1] for L in layers),
                "wo": cat(L["out"][0].view(L["out"][0].shape[0], -1) for L in layers), "bo": cat(L["out"][1] for L in layers),
                "w1": cat(L["up"][0].view(L["up"][0].shape[0], -1) for L in layers), "b1": cat(L["up"][1] for L in layers),
                "w2": cat(L["down"][0].view(L["down"][0].shape[0], -1) for L in layers), "b2": cat(L["down"][1] for L in layers),
                "ln1_g": cat(L["ln1"][0] for L in layers), "ln1_b": cat(L["ln1"][1] for L in layers),
                "ln2_g": cat(L["ln2"][0] for L in layers), "ln2_b": cat(L["ln2"][1] for L in layers),
                "lnf_g": f32(self.ln_f.weight), "lnf_b": f32(self.ln_f.bias),
                "eps1": float(layers[0]["ln1"][2]), "eps2": float(layers[0]["ln2"][2]), "epsf": float(self.ln_f.eps),
                "uniform_eps": all(L["ln1"][2] == layers[0]["ln1"][2] and L["ln2"][2] == layers[0]["ln2"][2] for L in layers),
            } if layers else None
            return {"layers": layers, "pos": f32(self.pos_emb), "stack": stack,
                    "lnf": (f32(self.ln_f.weight), f32(self.ln_f.bias), self.ln_f.eps)}
        return self._packed.get((_versions(*srcs), str(device)), build)

    def tokens(self, rgb, ir):
        """Everything up to and including ln_f: fp32 [B, 2*va*ha, d]."""
        if self.training and any(m.p > 0 for m in self.modules() if isinstance(m, nn.Dropout)):
            raise CftError("GPT in training mode: the embd / attn / resid dropouts (models/common.py:466-467,537,575) are not "
                           "implemented on the B200 forward path -- call model.eval()")
        rgb, ir = ops.to_nhwc_bf16(rgb), ops.to_nhwc_bf16(ir)
        assert rgb.shape == ir.shape
        b, c, _, _ = rgb.shape
        t = 2 * self.vert_anchors * self.horz_anchors
        wts = self._weights(rgb.device)
        x = ops.gpt_pool_tokens(rgb, ir, wts["pos"], self.vert_anchors, self.horz_anchors)   # fp32 [B,T,d]
        st = wts["stack"]
        if (self.fused_block and c <= self.fused_block_max_d and st is not None and st["uniform_eps"]
                and ops.gpt_block_supported(b, c, self.h, t)):
            # all layers + ln_f in ONE launch: a cluster of CTAs per image keeps the token tile on chip (csrc/cft_block.cu)
            return ops.gpt_block(x, st, self.h)
        x2d = x.view(b * t, c)
        for L in wts["layers"]:
            y = ops.layernorm(x2d, *L["ln1"])                                    # bf16 [B*T, d]
            qkv = ops.gemm(y, L["qkv"][0], L["qkv"][1])                         # bf16 [B*T, 3d]
            att = ops.attention(qkv, b, t, c, self.h)                            # bf16 [B*T, d]
            x2d = ops.gemm(att, L["out"][0], L["out"][1], residual=x2d, out_dtype=torch.float32)
            y = ops.layernorm(x2d, *L["ln2"])
            hid = ops.gemm(y, L["up"][0], L["up"][1], act=ACT_GELU)              # bf16 [B*T, 4d]
            x2d = ops.gemm(hid, L["down"][0], L["down"][1], residual=x2d, out_dtype=torch.float32)
        xf = ops.layernorm(x2d, *wts["lnf"], out_dtype=torch.float32)
        return xf.view(b, t, c)

    def forward(self, x):
        rgb, ir = x[0], x[1]
        tok = self.tokens(rgb, ir)
        h, w = rgb.shape[-2:]
        o_rgb, o_ir, _ = ops.gpt_unpool(tok, h, w, self.vert_anchors, self.horz_anchors)
        return o_rgb, o_ir

    def forward_fused(self, rgb, ir, out_rgb=None, out_ir=None, out_sum=None):
        """(rgb + up_rgb, ir + up_ir, their sum) in one un-pool pass (GPT + 2x Add2 + Add)."""
        rgb, ir = ops.to_nhwc_bf16(rgb), ops.to_nhwc_bf16(ir)
        tok = self.tokens(rgb, ir)
        h, w = rgb.shape[-2:]
        return ops.gpt_unpool(tok, h, w, self.vert_anchors, self.horz_anchors, x_rgb=rgb, x_ir=ir, want_sum=True,
                              out_rgb=out_rgb, out_ir=out_ir, out_sum=out_sum)


# ------------------------------------------------------------------------------ Detect
class Detect(nn.Module):
    """reference models/yolo_test.py:25-64. 1x1 head convs on the tcgen05 GEMM with fp32 output, then one
    kernel per level does view/permute (:48), sigmoid, grid/anchor decode (:54-57) and the cat (:59)
    in fp32.  Returns fp32 ``(z, [raw heads])`` in eval mode, the raw heads in training mode."""
    stride = None
    export = False

    def __init__(self, nc=80, anchors=(), ch=()):
        super().__init__()
        self.nc = nc
        self.no = nc + 5
        self.nl = len(anchors)
        self.na = len(anchors[0]) // 2
        self.grid = [torch.zeros(1)] * self.nl
        a = torch.tensor(anchors).float().view(self.nl, -1, 2)
        self.register_buffer('anchors', a)
        self.register_buffer('anchor_grid', a.clone().view(self.nl, 1, -1, 1, 1, 2))
        self.m = nn.ModuleList(nn.Conv2d(x, self.no * self.na, 1) for x in ch)
        self._packed = _Packed()

    def _weights(self, device):
        srcs = [p for p in self.m.parameters()] + [self.anchor_grid]

        def build():
            ws = [ops.pack_conv_weight(m.weight, m.bias, None, device=device) for m in self.m]
            ag = self.anchor_grid.detach().to(device=device, dtype=torch.float32).reshape(self.nl, self.na * 2).contiguous()
            return ws, ag
        return self._packed.get((_versions(*srcs), str(device)), build)

    def forward(self, x):
        if self.stride is None:
            raise CftError("Detect.stride is not set (Model.__init__ sets it, models/yolo_test.py:201)")
        xs = [ops.to_nhwc_bf16(t) for t in x]
        dev = xs[0].device
        ws, ag = self._weights(dev)
        b = xs[0].shape[0]
        rows = [self.na * t.shape[2] * t.shape[3] for t in xs]
        z = torch.empty((b, sum(rows), self.no), dtype=torch.float32, device=dev)
        raw, row0 = [], 0
        for i, t in enumerate(xs):
            _, c, ny, nx = t.shape
            w, bias = ws[i]
            _, ld = ops.nhwc_desc(t)
            a2d = t.as_strided((b * ny * nx, c), (ld, 1))                   # NHWC pixels as GEMM rows
            head = ops.gemm(a2d, w, bias, out_dtype=torch.float32)          # [B*ny*nx, pad8(na*no)] fp32
            raw.append(ops.detect_decode(head, b, ny, nx, self.na, self.no, float(self.stride[i]), ag[i], z, row0))
            row0 += rows[i]
        for i in range(self.nl):                                             # the reference overwrites the list (:46-48)
            try:
                x[i] = raw[i]
            except TypeError:
                pass
        return raw if self.training else (z, raw)
