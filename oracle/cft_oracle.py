"""CPU oracle for the two-stream yolov5-CFTx3 forward.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain fp32 ``torch.nn.functional`` calls on CPU tensors, the
algorithm of the reference forward path (DocF/multispectral-object-detection).  It is
NOT part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as
the checker / CPU baseline.  The product path (``multispectral-object-detection_b200``)
never imports anything from ``oracle/``.

Parity pinning: the reference ships no tests or golden vectors, and its
arithmetic lives in PyTorch (``requirements.txt``: torch>=1.7, unpinned; here 2.11.0).
The restatement is therefore pinned against the *unmodified reference itself*, imported
in the build container by ``oracle/ref_shim.py``; ``oracle/make_golden.py`` runs both on
the same seeded weights/inputs, asserts agreement, and commits the reference's outputs
under ``tests/golden/`` (the reference cannot travel to the GPU box).

Every function cites the reference file:line it follows (paths relative to the
reference root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

BN_EPS = 1e-3      # utils/torch_utils.py:149 (initialize_weights sets eps on every BatchNorm2d)
LN_EPS = 1e-5      # nn.LayerNorm default, models/common.py:529-530,572
STRIDES = (8.0, 16.0, 32.0)  # models/yolo_test.py:201 (hard-coded)


# ----------------------------------------------------------------------------------------
# graph construction: models/yolo_test.py:479-555 (parse_model)
# ----------------------------------------------------------------------------------------
def make_divisible(x, divisor):  # utils/general.py:210-212
    return math.ceil(x / divisor) * divisor


def build_spec(cfg: dict) -> Tuple[List[dict], List[int]]:
    """Resolve channels/depths of every layer row, as parse_model does.

    Returns (layers, save) where each layer is a dict
    ``{i, f, type, c1, c2, n, args}`` and ``save`` is the sorted save list
    (models/yolo_test.py:547,555).
    """
    anchors, nc, gd, gw = cfg["anchors"], cfg["nc"], cfg["depth_multiple"], cfg["width_multiple"]
    na = len(anchors[0]) // 2
    no = na * (nc + 5)
    ch: List[int] = [3]
    layers, save = [], []
    c2 = ch[-1]
    for i, (f, n, m, args) in enumerate(cfg["backbone"] + cfg["head"]):
        args = list(args)
        n = max(round(n * gd), 1) if n > 1 else n            # yolo_test.py:495
        spec = {"i": i, "f": f, "type": m, "n": 1, "args": None}
        if m in ("Conv", "SPP", "Focus", "C3"):
            if m == "Focus":
                c1, c2 = 3, args[0]                          # yolo_test.py:499-500
            else:
                c1, c2 = ch[f], args[0]
            if c2 != no:
                c2 = make_divisible(c2 * gw, 8)              # yolo_test.py:503,508
            rest = args[1:]
            if m == "C3":
                spec["n"] = n
                spec["shortcut"] = rest[0] if rest else True
            elif m == "SPP":
                spec["k"] = tuple(rest[0]) if rest else (5, 9, 13)
            else:  # Conv / Focus: (k, s)
                spec["k"] = rest[0] if len(rest) > 0 else 1
                spec["s"] = rest[1] if len(rest) > 1 else 1
            spec["c1"] = c1
        elif m == "Concat":
            c2 = sum(ch[x] for x in f)                       # yolo_test.py:517-518
        elif m in ("Add", "GPT"):
            c2 = ch[f[0]]                                    # yolo_test.py:519-530
        elif m == "Add2":
            c2 = ch[f[0]]
            spec["index"] = args[1]
        elif m == "Detect":
            spec["ch"] = [ch[x] for x in f]                  # yolo_test.py:531-534
            spec["nc"] = nc
            spec["anchors"] = anchors
        elif m == "nn.Upsample":
            c2 = ch[f]
            spec["scale"] = args[1]
        else:
            raise ValueError(f"module {m} is outside the x3 hot path")
        spec["c2"] = c2
        layers.append(spec)
        save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)  # :547
        if i == 0:
            ch = []
        ch.append(c2)
    return layers, sorted(set(save))


# ----------------------------------------------------------------------------------------
# modules: models/common.py
# ----------------------------------------------------------------------------------------
def conv_bn_silu(x: Tensor, sd: Dict[str, Tensor], p: str, k: int, s: int = 1) -> Tensor:
    """``Conv.forward``: SiLU(BN(conv2d(x))), pad=k//2, no conv bias (common.py:24-28,36-47).

    If the state dict was fused (``Model.fuse``, yolo_test.py:296-304: ``conv.bias`` present,
    no ``bn.*``), follows ``Conv.fuseforward`` (common.py:49-50).
    """
    w = sd[p + "conv.weight"]
    if p + "bn.weight" in sd:
        y = F.conv2d(x, w, None, stride=s, padding=k // 2)
        y = F.batch_norm(y, sd[p + "bn.running_mean"], sd[p + "bn.running_var"],
                         sd[p + "bn.weight"], sd[p + "bn.bias"], False, 0.0, BN_EPS)
    else:
        y = F.conv2d(x, w, sd[p + "conv.bias"], stride=s, padding=k // 2)
    return F.silu(y)


def focus(x: Tensor, sd, p: str, k: int) -> Tensor:
    """``Focus.forward`` (common.py:168-180): 2x2 space-to-depth, block-major channel order."""
    y = torch.cat([x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]], 1)
    return conv_bn_silu(y, sd, p + "conv.", k, 1)


def bottleneck(x: Tensor, sd, p: str, add: bool) -> Tensor:
    """``Bottleneck.forward`` (common.py:99-109) with e=1.0 as used inside C3 (:139)."""
    y = conv_bn_silu(conv_bn_silu(x, sd, p + "cv1.", 1), sd, p + "cv2.", 3)
    return x + y if add else y


def c3(x: Tensor, sd, p: str, n: int, shortcut: bool) -> Tensor:
    """``C3.forward`` (common.py:131-143): cv3(cat(m(cv1(x)), cv2(x)))."""
    a = conv_bn_silu(x, sd, p + "cv1.", 1)
    for j in range(n):
        a = bottleneck(a, sd, f"{p}m.{j}.", shortcut)     # c1 == c2 == c_ inside C3
    b = conv_bn_silu(x, sd, p + "cv2.", 1)
    return conv_bn_silu(torch.cat((a, b), 1), sd, p + "cv3.", 1)


def spp(x: Tensor, sd, p: str, ks: Sequence[int]) -> Tensor:
    """``SPP.forward`` (common.py:154-165): max-pools stride 1, pad k//2 (implicit -inf)."""
    x = conv_bn_silu(x, sd, p + "cv1.", 1)
    return conv_bn_silu(torch.cat([x] + [F.max_pool2d(x, k, 1, k // 2) for k in ks], 1), sd, p + "cv2.", 1)


def self_attention(x: Tensor, sd, p: str, h: int) -> Tensor:
    """``SelfAttention.forward`` (common.py:477-513); no mask / weights (never passed)."""
    b, n, d = x.shape
    dk = d // h
    q = F.linear(x, sd[p + "que_proj.weight"], sd[p + "que_proj.bias"]).view(b, n, h, dk).permute(0, 2, 1, 3)
    k = F.linear(x, sd[p + "key_proj.weight"], sd[p + "key_proj.bias"]).view(b, n, h, dk).permute(0, 2, 3, 1)
    v = F.linear(x, sd[p + "val_proj.weight"], sd[p + "val_proj.bias"]).view(b, n, h, dk).permute(0, 2, 1, 3)
    att = torch.matmul(q, k) / math.sqrt(dk)                 # :497
    att = torch.softmax(att, -1)                             # :506 (dropout = identity in eval)
    out = torch.matmul(att, v).permute(0, 2, 1, 3).contiguous().view(b, n, d)   # :510
    return F.linear(out, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])    # :511


def transformer_block(x: Tensor, sd, p: str, h: int) -> Tensor:
    """``myTransformerBlock.forward`` (common.py:539-546): pre-LN, erf-GELU MLP."""
    d = x.shape[-1]
    y = F.layer_norm(x, (d,), sd[p + "ln_input.weight"], sd[p + "ln_input.bias"], LN_EPS)
    x = x + self_attention(y, sd, p + "sa.", h)
    y = F.layer_norm(x, (d,), sd[p + "ln_output.weight"], sd[p + "ln_output.bias"], LN_EPS)
    y = F.linear(y, sd[p + "mlp.0.weight"], sd[p + "mlp.0.bias"])
    y = F.gelu(y)                                            # nn.GELU() default = erf form (:535)
    y = F.linear(y, sd[p + "mlp.2.weight"], sd[p + "mlp.2.bias"])
    return x + y


def gpt(rgb: Tensor, ir: Tensor, sd, p: str, h: int = 8, anchors_vh: Tuple[int, int] = (8, 8)):
    """``GPT.forward`` (common.py:593-639)."""
    bs, c, hh, ww = rgb.shape
    va, ha = anchors_vh
    r = F.adaptive_avg_pool2d(rgb, (va, ha))                 # :608
    i = F.adaptive_avg_pool2d(ir, (va, ha))                  # :609
    tok = torch.cat([r.view(bs, c, -1), i.view(bs, c, -1)], dim=2).permute(0, 2, 1).contiguous()  # :615-618
    x = sd[p + "pos_emb"] + tok                              # :621
    n_layer = 0
    while f"{p}trans_blocks.{n_layer}.ln_input.weight" in sd:
        n_layer += 1
    for l in range(n_layer):                                 # :622
        x = transformer_block(x, sd, f"{p}trans_blocks.{l}.", h)
    x = F.layer_norm(x, (c,), sd[p + "ln_f.weight"], sd[p + "ln_f.bias"], LN_EPS)   # :625
    x = x.view(bs, 2, va, ha, c).permute(0, 1, 4, 2, 3)      # :626-627
    r_out = x[:, 0].contiguous().view(bs, c, va, ha)
    i_out = x[:, 1].contiguous().view(bs, c, va, ha)
    r_out = F.interpolate(r_out, size=[hh, ww], mode="bilinear")    # :636 (align_corners=False)
    i_out = F.interpolate(i_out, size=[hh, ww], mode="bilinear")    # :637
    return r_out, i_out


def make_grid(nx: int, ny: int) -> Tensor:
    """``Detect._make_grid`` (yolo_test.py:61-64): grid[0,0,j,i] = (i, j)."""
    yv, xv = torch.meshgrid([torch.arange(ny), torch.arange(nx)], indexing="ij")
    return torch.stack((xv, yv), 2).view(1, 1, ny, nx, 2).float()


def detect(xs: List[Tensor], sd, p: str, nc: int, na: int = 3):
    """``Detect.forward`` in eval mode (yolo_test.py:41-59). Returns (z, [raw heads])."""
    no = nc + 5
    z, raw = [], []
    anchor_grid = sd[p + "anchor_grid"]
    for i, x in enumerate(xs):
        x = F.conv2d(x, sd[f"{p}m.{i}.weight"], sd[f"{p}m.{i}.bias"])          # :46
        bs, _, ny, nx = x.shape
        x = x.view(bs, na, no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()     # :48
        raw.append(x)
        y = x.sigmoid()
        grid = make_grid(nx, ny).to(y.device)
        xy = (y[..., 0:2] * 2.0 - 0.5 + grid) * STRIDES[i]                     # :55
        wh = (y[..., 2:4] * 2) ** 2 * anchor_grid[i]                           # :56
        y = torch.cat((xy, wh, y[..., 4:]), -1)
        z.append(y.view(bs, -1, no))                                            # :57
    return torch.cat(z, 1), raw                                                 # :59


def decode_heads(raw: List[Tensor], anchor_grid: Tensor) -> Tensor:
    """The decode half of Detect alone (yolo_test.py:54-59), for bit-exactness checks of
    grid/anchor indexing on identical raw heads."""
    z = []
    for i, x in enumerate(raw):
        bs, na, ny, nx, no = x.shape
        y = x.sigmoid()
        grid = make_grid(nx, ny).to(y.device)
        xy = (y[..., 0:2] * 2.0 - 0.5 + grid) * STRIDES[i]
        wh = (y[..., 2:4] * 2) ** 2 * anchor_grid[i]
        y = torch.cat((xy, wh, y[..., 4:]), -1)
        z.append(y.view(bs, -1, no))
    return torch.cat(z, 1)


# ----------------------------------------------------------------------------------------
# whole forward: models/yolo_test.py:235-272 (forward_once)
# ----------------------------------------------------------------------------------------
@torch.no_grad()
def forward(sd: Dict[str, Tensor], cfg: dict, x_rgb: Tensor, x_ir: Tensor, capture: bool = False):
    """Eval forward of the two-stream model. Returns ``(z, [raw P3, P4, P5])`` and, with
    ``capture=True``, also the list of every layer's output (layer index -> tensor/tuple)."""
    layers, save = build_spec(cfg)
    y: List = []
    outs: List = []
    x = x_rgb
    for L in layers:
        i, f, t = L["i"], L["f"], L["type"]
        p = f"model.{i}."
        if f != -1 and f != -4:                                               # :247-250
            x = y[f] if isinstance(f, int) else [x if j == -1 else y[j] for j in f]
        if f == -4:
            x = x_ir                                                          # :262-263
        if t == "Conv":
            x = conv_bn_silu(x, sd, p, L["k"], L["s"])
        elif t == "Focus":
            x = focus(x, sd, p, L["k"])
        elif t == "C3":
            x = c3(x, sd, p, L["n"], L["shortcut"])
        elif t == "SPP":
            x = spp(x, sd, p, L["k"])
        elif t == "GPT":
            x = gpt(x[0], x[1], sd, p)
        elif t == "Add2":
            x = x[0] + x[1][L["index"]]                                       # common.py:239-242
        elif t == "Add":
            x = x[0] + x[1]                                                   # common.py:229
        elif t == "Concat":
            x = torch.cat(x, 1)                                               # common.py:219
        elif t == "nn.Upsample":
            x = F.interpolate(x, scale_factor=L["scale"], mode="nearest")
        elif t == "Detect":
            x = detect(list(x), sd, p, L["nc"])
        y.append(x if i in save else None)                                    # :266
        if capture:
            outs.append(x)
    return (x[0], x[1], outs) if capture else x


# ----------------------------------------------------------------------------------------
# seeded synthetic weights with the reference's state_dict keys
# ----------------------------------------------------------------------------------------
def init_state(cfg: dict, seed: int = 0, n_layer: int = 8, tokens: int = 128, gain: float = 4.0, det_gain: float = 1.0) -> Dict[str, Tensor]:
    """Seeded "randomised-stats" weights under the reference's
    state_dict keys: conv ~ U(+-1/sqrt(fan_in)) scaled so activations stay O(1) through the
    SiLU stack, BN gamma~U(.5,1.5) beta~N(0,.1) mean~N(0,.1) var~U(.5,1.5), Linear N(0,.02)
    (common.py:583-591) with small random biases, pos_emb~N(0,.02), LN gamma~U(.5,1.5).
    The generator is a CPU ``torch.Generator`` so the values are identical on any host."""
    g = torch.Generator().manual_seed(seed)
    layers, _ = build_spec(cfg)
    sd: Dict[str, Tensor] = {}

    def conv(p, c1, c2, k):
        fan_in = c1 * k * k
        bound = math.sqrt(gain / fan_in)                  # keeps variance ~ through SiLU stacks
        sd[p + "conv.weight"] = (torch.rand(c2, c1, k, k, generator=g) * 2 - 1) * bound
        sd[p + "bn.weight"] = torch.rand(c2, generator=g) + 0.5
        sd[p + "bn.bias"] = torch.randn(c2, generator=g) * 0.1
        sd[p + "bn.running_mean"] = torch.randn(c2, generator=g) * 0.1
        sd[p + "bn.running_var"] = torch.rand(c2, generator=g) + 0.5
        sd[p + "bn.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    def linear(p, cin, cout):
        sd[p + "weight"] = torch.randn(cout, cin, generator=g) * 0.02
        sd[p + "bias"] = torch.randn(cout, generator=g) * 0.02

    def ln(p, d):
        sd[p + "weight"] = torch.rand(d, generator=g) + 0.5
        sd[p + "bias"] = torch.randn(d, generator=g) * 0.1

    for L in layers:
        i, t = L["i"], L["type"]
        p = f"model.{i}."
        if t == "Conv":
            conv(p, L["c1"], L["c2"], L["k"])
        elif t == "Focus":
            conv(p + "conv.", 12, L["c2"], L["k"])
        elif t == "C3":
            c_ = int(L["c2"] * 0.5)
            conv(p + "cv1.", L["c1"], c_, 1)
            conv(p + "cv2.", L["c1"], c_, 1)
            conv(p + "cv3.", 2 * c_, L["c2"], 1)
            for j in range(L["n"]):
                conv(f"{p}m.{j}.cv1.", c_, c_, 1)
                conv(f"{p}m.{j}.cv2.", c_, c_, 3)
        elif t == "SPP":
            c_ = L["c1"] // 2
            conv(p + "cv1.", L["c1"], c_, 1)
            conv(p + "cv2.", c_ * (len(L["k"]) + 1), L["c2"], 1)
        elif t == "GPT":
            d = L["c2"]
            sd[p + "pos_emb"] = torch.randn(1, tokens, d, generator=g) * 0.02
            for l in range(n_layer):
                q = f"{p}trans_blocks.{l}."
                ln(q + "ln_input.", d)
                ln(q + "ln_output.", d)
                for nm in ("que_proj", "key_proj", "val_proj", "out_proj"):
                    linear(f"{q}sa.{nm}.", d, d)
                linear(q + "mlp.0.", d, 4 * d)
                linear(q + "mlp.2.", 4 * d, d)
            ln(p + "ln_f.", d)
        elif t == "Detect":
            a = torch.tensor(L["anchors"]).float().view(3, -1, 2)
            sd[p + "anchors"] = a / torch.tensor(STRIDES).view(-1, 1, 1)     # yolo_test.py:203
            sd[p + "anchor_grid"] = a.clone().view(3, 1, -1, 1, 1, 2)        # yolo_test.py:38
            no = (L["nc"] + 5) * 3
            for j, c in enumerate(L["ch"]):
                sd[f"{p}m.{j}.weight"] = (torch.rand(no, c, 1, 1, generator=g) * 2 - 1) * (det_gain / math.sqrt(c))
                sd[f"{p}m.{j}.bias"] = torch.randn(no, generator=g) * 0.5
    return sd


def make_inputs(batch: int, height: int, width: int, seed: int = 1):
    """Synthetic image pair in [0,1): RGB then IR from one CPU generator."""
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(batch, 3, height, width, generator=g),
            torch.rand(batch, 3, height, width, generator=g))


def conv_linear_flops(cfg: dict, height: int, width: int) -> float:
    """Algorithmic FLOPs per pair (2*MAC over conv + linear + attention core), SURVEY.md §8(d)."""
    layers, _ = build_spec(cfg)
    total = 0.0
    shapes: Dict[int, Tuple[int, int]] = {}

    def cflops(h, w, c1, c2, k):
        return 2.0 * h * w * c1 * c2 * k * k

    cur = (height, width)
    for L in layers:
        i, f, t = L["i"], L["f"], L["type"]
        src = lambda idx: shapes[i - 1] if idx == -1 else shapes[idx]
        if f == -4 or i == 0:
            cur = (height, width)
        elif isinstance(f, int):
            cur = src(f)
        else:
            cur = src(f[0])
        h, w = cur
        if t == "Focus":
            h, w = h // 2, w // 2
            total += cflops(h, w, 12, L["c2"], L["k"])
        elif t == "Conv":
            h, w = (h + L["s"] - 1) // L["s"], (w + L["s"] - 1) // L["s"]
            total += cflops(h, w, L["c1"], L["c2"], L["k"])
        elif t == "C3":
            c_ = L["c2"] // 2
            total += 2 * cflops(h, w, L["c1"], c_, 1) + cflops(h, w, 2 * c_, L["c2"], 1)
            total += L["n"] * (cflops(h, w, c_, c_, 1) + cflops(h, w, c_, c_, 3))
        elif t == "SPP":
            c_ = L["c1"] // 2
            total += cflops(h, w, L["c1"], c_, 1) + cflops(h, w, 4 * c_, L["c2"], 1)
        elif t == "GPT":
            d = L["c2"]
            total += 8 * (2.0 * 128 * (4 * d * d + 8 * d * d)) + 8 * 4.0 * 128 * 128 * d
        elif t == "nn.Upsample":
            h, w = h * 2, w * 2
        elif t == "Detect":
            for j, c in enumerate(L["ch"]):
                hh, ww = shapes[L["f"][j]]
                total += cflops(hh, ww, c, (L["nc"] + 5) * 3, 1)
        cur = (h, w)
        shapes[i] = cur
    return total
