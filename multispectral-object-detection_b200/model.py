"""Host-side mirror of the reference's two-stream ``Model`` (models/yolo_test.py:165-327) and of
``parse_model`` (:479-555), plus the two ways to drop the B200 modules into the reference itself:

* ``Model(cfg)``            -- stand-alone: builds the graph from a dict (``config.x3_config``) or a
                               reference yaml path; same layer protocol (``m.i/m.f/m.type/m.np``),
                               same ``state_dict`` keys, same ``forward(x_rgb, x_ir)`` return values.
* ``install(ref_yolo_test)`` -- rebinds ``Conv, Focus, Bottleneck, C3, SPP, Concat, Add, Add2, GPT,
                               Detect`` (and ``nn.Upsample``) in the reference's own namespace so its
                               unmodified ``Model``/``parse_model``/``forward_once`` build and run
                               the B200 modules (the ``eval(m)`` boundary, models/yolo_test.py:488).
* ``convert(ref_model)``     -- swaps the modules of an already built / unpickled reference ``Model``
                               (checkpoints are pickled modules, train.py:850-857).

On top of the per-layer walk of ``forward_once`` (:235-272) the stand-alone ``Model`` plans buffers:
producers of a ``Concat`` write straight into channel slices of its buffer, and every
``GPT -> Add2, Add2 -> Add`` group of the x3 graphs runs as one fused un-pool pass.
"""
from __future__ import annotations

import logging
import math
import os
from copy import deepcopy
from fractions import Fraction
from pathlib import Path
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import modules as M
from . import ops
from ._lib import CftError
from .config import make_divisible

logger = logging.getLogger(__name__)

REGISTRY = {
    "Conv": M.Conv, "Focus": M.Focus, "Bottleneck": M.Bottleneck, "C3": M.C3, "SPP": M.SPP, "Concat": M.Concat,
    "Add": M.Add, "Add2": M.Add2, "GPT": M.GPT, "Detect": M.Detect, "nn.Upsample": M.Upsample,
}
REBOUND_NAMES = ("Conv", "Focus", "Bottleneck", "C3", "SPP", "Concat", "Add", "Add2", "GPT", "Detect")


def parse_model(d: dict, ch: List[int]):
    """Mirror of reference models/yolo_test.py:479-555 restricted to the hot-path module set."""
    anchors, nc, gd, gw = d['anchors'], d['nc'], d['depth_multiple'], d['width_multiple']
    na = (len(anchors[0]) // 2) if isinstance(anchors, list) else anchors
    no = na * (nc + 5)
    layers, save, c2 = [], [], ch[-1]
    scope = {"nc": nc, "anchors": anchors, "None": None, "False": False, "True": True}
    for i, (f, n, mname, args) in enumerate(d['backbone'] + d['head']):
        if mname not in REGISTRY:
            raise CftError(f"module '{mname}' (layer {i}) is outside the CFTx3 hot path")
        m = REGISTRY[mname]
        args = [scope.get(a, a) if isinstance(a, str) else a for a in args]          # :489-493
        n = max(round(n * gd), 1) if n > 1 else n                                     # :495
        if m in (M.Conv, M.Bottleneck, M.SPP, M.Focus, M.C3):
            if m is M.Focus:
                c1, c2 = 3, args[0]                                                   # :499-500
            else:
                c1, c2 = ch[f], args[0]
            if c2 != no:
                c2 = make_divisible(c2 * gw, 8)                                       # :503,508
            args = [c1, c2, *args[1:]]
            if m is M.C3:
                args.insert(2, n)                                                     # :511-513
                n = 1
        elif m is M.Concat:
            c2 = sum(ch[x] for x in f)
        elif m is M.Add:
            c2 = ch[f[0]]
            args = [c2]
        elif m is M.Add2:
            c2 = ch[f[0]]
            args = [c2, args[1]]
        elif m is M.GPT:
            c2 = ch[f[0]]
            args = [c2]
        elif m is M.Detect:
            args.append([ch[x] for x in f])
            if isinstance(args[1], int):
                args[1] = [list(range(args[1] * 2))] * len(f)
        else:
            c2 = ch[f]
        m_ = nn.Sequential(*[m(*args) for _ in range(n)]) if n > 1 else m(*args)     # :542
        t = mname if mname != "nn.Upsample" else "torch.nn.modules.upsampling.Upsample"
        npar = sum(x.numel() for x in m_.parameters())
        m_.i, m_.f, m_.type, m_.np = i, f, t, npar                                    # :545
        save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)    # :547
        layers.append(m_)
        if i == 0:
            ch = []
        ch.append(c2)
    return nn.Sequential(*layers), sorted(save)


def check_anchor_order(m):  # reference utils/autoanchor.py:12-20
    a = m.anchor_grid.prod(-1).view(-1)
    da = a[-1] - a[0]
    ds = m.stride[-1] - m.stride[0]
    if da.sign() != ds.sign():
        m.anchors[:] = m.anchors.flip(0)
        m.anchor_grid[:] = m.anchor_grid.flip(0)


class Model(nn.Module):
    """Two-stream detector (reference models/yolo_test.py:165-327), forward path on libcft_b200."""

    def __init__(self, cfg, ch=3, nc=None, anchors=None):
        super().__init__()
        if isinstance(cfg, dict):
            self.yaml = deepcopy(cfg)
        else:
            import yaml
            self.yaml_file = Path(cfg).name
            with open(cfg) as f:
                self.yaml = yaml.safe_load(f)
        ch = self.yaml['ch'] = self.yaml.get('ch', ch)
        if nc and nc != self.yaml['nc']:
            self.yaml['nc'] = nc
        if anchors:
            self.yaml['anchors'] = round(anchors)
        self.model, self.save = parse_model(deepcopy(self.yaml), ch=[ch])
        self.names = [str(i) for i in range(self.yaml['nc'])]
        m = self.model[-1]
        if isinstance(m, M.Detect):
            m.stride = torch.Tensor([8.0, 16.0, 32.0])                                # :201 (hard-coded)
            m.anchors /= m.stride.view(-1, 1, 1)                                      # :203
            check_anchor_order(m)
            self.stride = m.stride
            self._initialize_biases()
        for mod in self.modules():                                                    # utils/torch_utils.py:144-153
            if type(mod) is nn.BatchNorm2d:
                mod.eps = 1e-3
                mod.momentum = 0.03
        self._plan = _plan_graph(self.model)

    # -------------------------------------------------------------------------------- forward
    def forward(self, x, x2, augment=False, profile=False):
        if augment:
            raise CftError("augment=True is broken for the two-stream reference as well "
                           "(models/yolo_test.py:222 drops x2); not on the hot path")
        return self.forward_once(x, x2, profile)

    # run the IR branch between fusion points on a side CUDA stream (fills the tail waves); CFT_ONE_STREAM=1 disables
    two_streams = os.environ.get("CFT_ONE_STREAM") is None

    def forward_once(self, x, x2, profile=False):
        """Layer walk of reference models/yolo_test.py:235-272 with buffer planning on top.

        The RGB and IR chains between two CFT blocks are independent; the IR chain is issued on a side stream
        (fork after its source layer, join at the GPT) so that its persistent kernels fill the SMs left idle by
        the tail wave of the RGB chain's kernels (and vice versa).  Fork/join uses CUDA events only, so the walk is
        capturable into a CUDA graph."""
        plan = self._plan
        cap = getattr(self, "_capture", None)              # dict: layer index -> output, filled when set (tests)
        y: List = []
        concat_bufs: Dict[int, torch.Tensor] = {}
        fused: Dict[int, torch.Tensor] = {}
        chains = plan["ir_chains"] if (self.two_streams and x.is_cuda) else {}
        main = torch.cuda.current_stream() if chains else None
        side = self._side_stream(x.device) if chains else None
        src_events: Dict[int, torch.cuda.Event] = {}
        wanted_src = {src for (_, src) in chains.values()}
        chain_end, side_active = -1, False
        if chains and -4 in wanted_src:
            src_events[-4] = main.record_event()
        for m in self.model:
            i = m.i
            if m.f != -1 and m.f != -4:
                x = y[m.f] if isinstance(m.f, int) else [x if j == -1 else y[j] for j in m.f]
            if m.f == -4:
                x = x2
            if i in chains:                                    # fork: the IR chain starts here
                chain_end, src = chains[i]
                side.wait_event(src_events[src])
                side_active = True
            if side_active and isinstance(m, M.GPT):           # join at the fusion block
                main.wait_stream(side)
                side_active = False
            on_side = side_active and i <= chain_end
            ctx = torch.cuda.stream(side) if on_side else _NullCtx()
            with ctx:
                if i in fused:                                 # Add2 / Add already produced by the fused GPT pass
                    x = fused.pop(i)
                elif i in plan["gpt_groups"]:
                    g = plan["gpt_groups"][i]
                    outs = {}
                    for key in ("rgb", "ir", "sum"):
                        outs[key] = self._slot(plan, concat_bufs, g[key], x[0]) if g[key] in plan["slots"] else None
                    o_rgb, o_ir, o_sum = m.forward_fused(x[0], x[1], out_rgb=outs["rgb"], out_ir=outs["ir"], out_sum=outs["sum"])
                    fused[g["rgb"]], fused[g["ir"]], fused[g["sum"]] = o_rgb, o_ir, o_sum
                    if chains:
                        for key in ("rgb", "ir"):              # these feed the next RGB / IR chains
                            if g[key] in wanted_src:
                                src_events[g[key]] = main.record_event()
                    x = None                                   # the raw GPT tuple is never materialised
                elif i in plan["slots"]:
                    ref = x[0] if isinstance(x, (list, tuple)) else x
                    x = m(x, out=self._slot(plan, concat_bufs, i, ref, m))
                else:
                    x = m(x)
            if chains and not on_side and i in wanted_src and i not in src_events:
                src_events[i] = main.record_event()
            if cap is not None:                                # debug / tests: every layer's output (None for a fused GPT)
                cap[i] = x
            y.append(x if i in self.save else None)
        if side_active:                                        # graph without a closing GPT (not the x3 layout)
            main.wait_stream(side)
        return x

    def _side_stream(self, device):
        st = getattr(self, "_side", None)
        if st is None or st.device != device:
            st = torch.cuda.Stream(device)
            object.__setattr__(self, "_side", st)
        return st

    def _slot(self, plan, bufs, i, ref, m=None):
        """Channel-slice view of the Concat buffer that layer i's output is planned into.  ``ref`` is the producer's (first)
        input; its resolution times the producer's planned scale (``_plan_graph``: Focus / strided Conv shrink, Upsample
        doubles) is the resolution of the buffer."""
        cat_i, c0, c1, ctot, (mul, div) = plan["slots"][i]
        key = cat_i
        if key not in bufs:
            h, w = (ref.shape[2] * mul + div - 1) // div, (ref.shape[3] * mul + div - 1) // div
            bufs[key] = ops.empty_nhwc(ref.shape[0], ctot, h, w, ref.device)
        return M.concat_slot(bufs[key], c0, c1)

    # -------------------------------------------------------------------------------- misc API
    def _initialize_biases(self, cf=None):  # reference models/yolo_test.py:274-282
        m = self.model[-1]
        for mi, s in zip(m.m, m.stride):
            b = mi.bias.view(m.na, -1)
            b.data[:, 4] += math.log(8 / (640 / s) ** 2)
            b.data[:, 5:] += math.log(0.6 / (m.nc - 0.99)) if cf is None else torch.log(cf / cf.sum())
            mi.bias = torch.nn.Parameter(b.view(-1), requires_grad=True)

    def fuse(self):
        """reference models/yolo_test.py:296-304: fold BN into every Conv (kept for API parity;
        the kernels always run on folded weights, fused or not)."""
        for m in self.modules():
            if type(m) is M.Conv and hasattr(m, 'bn'):
                m.conv = fuse_conv_and_bn(m.conv, m.bn)
                delattr(m, 'bn')
        return self

    def info(self, verbose=False, img_size=640):
        n_p = sum(x.numel() for x in self.parameters())
        logger.info(f"Model Summary: {len(list(self.modules()))} layers, {n_p} parameters")
        return n_p


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def fuse_conv_and_bn(conv, bn):
    """reference utils/torch_utils.py:181-201"""
    fused = nn.Conv2d(conv.in_channels, conv.out_channels, kernel_size=conv.kernel_size, stride=conv.stride,
                      padding=conv.padding, groups=conv.groups, bias=True).requires_grad_(False).to(conv.weight.device)
    w_conv = conv.weight.clone().view(conv.out_channels, -1)
    w_bn = torch.diag(bn.weight.div(torch.sqrt(bn.eps + bn.running_var)))
    fused.weight.copy_(torch.mm(w_bn, w_conv).view(fused.weight.shape))
    b_conv = torch.zeros(conv.weight.size(0), device=conv.weight.device) if conv.bias is None else conv.bias
    b_bn = bn.bias - bn.weight.mul(bn.running_mean).div(torch.sqrt(bn.running_var + bn.eps))
    fused.bias.copy_(torch.mm(w_bn, b_conv.reshape(-1, 1)).reshape(-1) + b_bn)
    return fused


# ------------------------------------------------------------------------------------ planning
def _plan_graph(layers: nn.Sequential) -> dict:
    """Static analysis of the layer list.

    slots[i] = (concat layer, c0, c1, c_total, (mul, div)) when layer i's output can be written directly into
    the buffer of a later Concat (i must be a single-tensor producer that accepts ``out=``; its output is
    ``ceil(input * mul / div)`` pixels high / wide; all producers of one Concat must sit at the same total stride);
    gpt_groups[g] = {rgb: Add2 idx, ir: Add2 idx, sum: Add idx} for every GPT whose two outputs are
    consumed only by an Add2 pair that is merged by one Add (the x3 pattern, yaml rows 10-12 + 29).
    """
    n = len(layers)
    out_ch: Dict[int, int] = {}
    consumers: Dict[int, List[int]] = {i: [] for i in range(n)}
    for m in layers:
        srcs = [m.f] if isinstance(m.f, int) else list(m.f)
        for s in srcs:
            if s == -4 or (s == -1 and m.i == 0):
                continue
            consumers[(m.i - 1) if s == -1 else s].append(m.i)

    def channels(m):
        if isinstance(m, M.Conv):
            return m.conv.out_channels
        if isinstance(m, M.C3):
            return m.cv3.conv.out_channels
        if isinstance(m, M.SPP):
            return m.cv2.conv.out_channels
        if isinstance(m, M.Focus):
            return m.conv.conv.out_channels
        return None

    for m in layers:
        c = channels(m)
        srcs = [m.f] if isinstance(m.f, int) else list(m.f)
        srcs = [(m.i - 1) if s == -1 else s for s in srcs if s != -4 and not (s == -1 and m.i == 0)]
        if c is None:
            if isinstance(m, M.Concat):
                c = sum(out_ch[s] for s in srcs)
            elif isinstance(m, M.Detect):
                c = 0
            else:
                c = out_ch[srcs[0]]
        out_ch[m.i] = c

    gpt_groups = {}
    for m in layers:
        if not isinstance(m, M.GPT):
            continue
        g = m.i
        cons = consumers[g]
        if len(cons) != 2 or not all(isinstance(layers[c], M.Add2) for c in cons):
            continue
        a, b = (layers[c] for c in cons)
        if {a.index, b.index} != {0, 1}:
            continue
        rgb_l, ir_l = (a, b) if a.index == 0 else (b, a)
        # Add2 inputs must be [stream feature, gpt] with the stream feature = the GPT's own inputs
        if list(rgb_l.f) != [m.f[0], g] or list(ir_l.f) != [m.f[1], g]:
            continue
        adds = [c for c in consumers[rgb_l.i] if isinstance(layers[c], M.Add) and set(layers[c].f) == {rgb_l.i, ir_l.i}]
        if len(adds) != 1:
            continue
        gpt_groups[g] = {"rgb": rgb_l.i, "ir": ir_l.i, "sum": adds[0]}

    # per-layer scale relative to the layer's own (first) input, and the total stride relative to the image
    rel: Dict[int, tuple] = {}
    total: Dict[int, Fraction] = {}
    for m in layers:
        if isinstance(m, M.Focus):
            r = (1, 2)
        elif isinstance(m, M.Conv):
            r = (1, int(m.conv.stride[0]))
        elif isinstance(m, M.Upsample):
            r = (2, 1)
        else:
            r = (1, 1)
        rel[m.i] = r
        srcs = [m.f] if isinstance(m.f, int) else list(m.f)
        first = srcs[0]
        base = Fraction(1) if (first == -4 or (first == -1 and m.i == 0)) else total[(m.i - 1) if first == -1 else first]
        total[m.i] = base * Fraction(r[1], r[0])

    writable = (M.Conv, M.C3, M.SPP, M.Focus, M.Upsample, M.Add, M.Add2)
    slots = {}
    for m in layers:
        if not isinstance(m, M.Concat):
            continue
        srcs = [(m.i - 1) if s == -1 else s for s in m.f]
        ctot = sum(out_ch[s] for s in srcs)
        c0 = 0
        ok = all(isinstance(layers[s], writable) and s not in slots for s in srcs) and len(set(srcs)) == len(srcs)
        ok = ok and len({total[s] for s in srcs}) == 1          # else: plain Concat copies (and its own shape error)
        for s in srcs:
            if ok:
                slots[s] = (m.i, c0, c0 + out_ch[s], ctot, rel[s])
            c0 += out_ch[s]
    # Two-stream execution: the IR branch between two fusion points is a chain of single-input layers that does not
    # depend on the RGB chain listed just before it in the yaml (rows 5-9, 15-16, 23-25).  chains[first] = (last, src):
    # layers first..last may run on a side CUDA stream once layer ``src`` (or the IR image, src = -4) is available;
    # the GPT that consumes ``last`` joins the streams.
    chains = {}
    for m in layers:
        if not isinstance(m, M.GPT):
            continue
        ia = m.f[1]
        first = ia
        while isinstance(layers[first].f, int) and layers[first].f == -1 and first > 0:
            first -= 1
        src = layers[first].f
        if not isinstance(src, int) or first <= m.f[0]:
            continue                                       # not the "RGB chain, then IR chain" layout
        if all(isinstance(layers[i], (M.Conv, M.C3, M.SPP, M.Focus)) for i in range(first, ia + 1)):
            chains[first] = (ia, src)
    return {"slots": slots, "gpt_groups": gpt_groups, "out_ch": out_ch, "ir_chains": chains}


# ------------------------------------------------------------------------------------ drop-in
def install(ref_module) -> dict:
    """Rebind the hot-path class names inside the reference's ``models.yolo_test`` namespace (and
    ``models.common`` if given the package) so the reference's unmodified ``Model(cfg)`` builds B200
    modules.  Returns the previous bindings (pass to ``uninstall``)."""
    prev = {}
    for name in REBOUND_NAMES:
        prev[name] = getattr(ref_module, name, None)
        setattr(ref_module, name, REGISTRY[name])
    ref_nn = getattr(ref_module, "nn", None)
    if ref_nn is not None:
        # ``nn.Upsample`` is resolved as eval("nn.Upsample"): give the reference namespace an ``nn`` proxy
        proxy = _NNProxy(ref_nn)
        prev["nn"] = ref_nn
        setattr(ref_module, "nn", proxy)
    return prev


def uninstall(ref_module, prev: dict):
    for name, val in prev.items():
        if val is not None:
            setattr(ref_module, name, val)


class _NNProxy:
    """``torch.nn`` with ``Upsample`` replaced by the B200 module (everything else passes through)."""

    def __init__(self, real):
        self._real = real

    def __getattr__(self, k):
        if k == "Upsample":
            return M.Upsample
        return getattr(self._real, k)


def convert(ref_model: nn.Module) -> nn.Module:
    """Swap the modules of a built reference two-stream ``Model`` for B200 ones, keeping its weights.
    Works on fused (``conv.bias``, no ``bn``) and unfused checkpoints."""
    new_layers = []
    for m in ref_model.model:
        new_layers.append(_convert_module(m))
        for attr in ("i", "f", "type", "np"):
            setattr(new_layers[-1], attr, getattr(m, attr))
    ref_model.model = nn.Sequential(*new_layers)
    return ref_model


def _convert_module(m: nn.Module) -> nn.Module:
    name = type(m).__name__
    if name == "Upsample":
        return M.Upsample(m.size, m.scale_factor, m.mode).train(m.training)
    if name not in REGISTRY:
        raise CftError(f"convert: module {name} is outside the CFTx3 hot path")
    if name == "Conv":
        c = m.conv
        new = M.Conv(c.in_channels, c.out_channels, c.kernel_size[0], c.stride[0], c.padding[0], c.groups,
                     True if isinstance(m.act, nn.SiLU) else m.act)
        if not hasattr(m, "bn"):
            new.conv = nn.Conv2d(c.in_channels, c.out_channels, c.kernel_size, c.stride, c.padding, groups=c.groups, bias=True)
            delattr(new, "bn")
        else:
            # load_state_dict copies weights and running statistics only: the fold needs the source's eps as well
            # (initialize_weights sets 1e-3, utils/torch_utils.py:144-153; nn.BatchNorm2d's default is 1e-5)
            new.bn.eps, new.bn.momentum = m.bn.eps, m.bn.momentum
    elif name == "Focus":
        c = m.conv.conv
        new = M.Focus(c.in_channels // 4, c.out_channels, c.kernel_size[0], c.stride[0], c.padding[0], c.groups)
        new.conv = _convert_module(m.conv)
    elif name == "Bottleneck":
        new = M.Bottleneck(m.cv1.conv.in_channels, m.cv2.conv.out_channels, m.add, 1, 1.0)
        new.cv1, new.cv2, new.add = _convert_module(m.cv1), _convert_module(m.cv2), m.add
    elif name == "C3":
        new = M.C3(m.cv1.conv.in_channels, m.cv3.conv.out_channels, 0)
        new.cv1, new.cv2, new.cv3 = (_convert_module(x) for x in (m.cv1, m.cv2, m.cv3))
        new.m = nn.Sequential(*[_convert_module(b) for b in m.m])
    elif name == "SPP":
        ks = tuple(p.kernel_size for p in m.m)
        new = M.SPP(m.cv1.conv.in_channels, m.cv2.conv.out_channels, ks)
        new.cv1, new.cv2 = _convert_module(m.cv1), _convert_module(m.cv2)
    elif name == "Concat":
        new = M.Concat(m.d)
    elif name == "Add":
        new = M.Add(m.arg)
    elif name == "Add2":
        new = M.Add2(0, m.index)
    elif name == "GPT":
        new = M.GPT(m.n_embd, h=m.trans_blocks[0].sa.h, n_layer=len(m.trans_blocks), vert_anchors=m.vert_anchors,
                    horz_anchors=m.horz_anchors)
    elif name == "Detect":
        anchors = m.anchor_grid.view(m.nl, -1).tolist()
        new = M.Detect(m.nc, anchors, [c.in_channels for c in m.m])
        new.stride = m.stride
    if name in ("Conv", "GPT", "Detect"):
        new.load_state_dict(m.state_dict(), strict=True)
    new.train(m.training)                       # fresh modules start in training mode: keep the source's mode (eval checkpoints)
    return new
