"""CPU restatement of the reference's post-processing step ``non_max_suppression`` -- TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module; the product
path (``multispectral-object-detection_b200.nms``) never does.

Follows reference ``utils/general.py:455-544`` (``non_max_suppression``), ``:299-306`` (``xywh2xyxy``) and the greedy
suppression of ``torchvision.ops.nms`` (``utils/general.py:527``).  torchvision is a third-party dependency of the
reference (``requirements.txt``: ``torchvision>=0.8.1``, unpinned; 0.26.0 in this image); its published CPU algorithm
(``torchvision/csrc/ops/cpu/nms_kernel.cpp``) is restated in ``nms_greedy``:

    areas = (x2 - x1) * (y2 - y1); order = stable argsort(scores, descending)
    for i in order (not yet suppressed): keep i; for every later, not yet suppressed j:
        inter = max(0, min(x2_i, x2_j) - max(x1_i, x1_j)) * max(0, min(y2_i, y2_j) - max(y1_i, y1_j))
        suppress j  iff  inter / (area_i + area_j - inter) > iou_thres          (all fp32, this operation order)

Pinned by ``oracle/make_golden_nms.py`` / ``tests/test_nms_cpu.py`` against the unmodified reference function
(which calls the installed torchvision) on seeded inputs: identical outputs, bit for bit.

Not restated: the ``labels`` (autolabelling) and ``merge`` branches (``merge = False`` is hard-coded at ``:470``) and
the wall-clock ``time_limit`` break (``:539-541``) -- none is deterministic / on the inference path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

MAX_WH = 4096        # utils/general.py:464
MAX_DET = 300        # :465
MAX_NMS = 30000      # :466


def xywh2xyxy(x: np.ndarray) -> np.ndarray:
    """utils/general.py:299-306 (fp32: ``x - w / 2``, ``x + w / 2``)."""
    y = np.empty_like(x)
    two = np.float32(2)
    y[:, 0] = x[:, 0] - x[:, 2] / two
    y[:, 1] = x[:, 1] - x[:, 3] / two
    y[:, 2] = x[:, 0] + x[:, 2] / two
    y[:, 3] = x[:, 1] + x[:, 3] / two
    return y


def nms_greedy(boxes: np.ndarray, scores: np.ndarray, iou_thres: float, limit: Optional[int] = None) -> np.ndarray:
    """torchvision.ops.nms (CPU kernel) in fp32 numpy; returns kept indices in descending score order.
    ``limit`` stops after that many kept boxes (the reference truncates to max_det afterwards, ``:528-529`` --
    identical, because a greedy decision depends on earlier kept boxes only)."""
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), np.int64)
    boxes = boxes.astype(np.float32, copy=False)
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    order = np.argsort(-scores.astype(np.float32), kind="stable")       # stable, descending
    thr = np.float32(iou_thres)
    zero = np.float32(0)
    suppressed = np.zeros(n, bool)
    keep: List[int] = []
    for pos in range(n):
        i = order[pos]
        if suppressed[i]:
            continue
        keep.append(int(i))
        if limit is not None and len(keep) >= limit:
            break
        rest = order[pos + 1:]
        rest = rest[~suppressed[rest]]
        if rest.size == 0:
            continue
        w = np.maximum(zero, np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]))
        h = np.maximum(zero, np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]))
        inter = w * h
        ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr > thr]] = True
    return np.asarray(keep, np.int64)


def non_max_suppression(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45,
                        classes: Optional[Sequence[int]] = None, agnostic: bool = False, multi_label: bool = False,
                        max_det: int = MAX_DET) -> List[torch.Tensor]:
    """utils/general.py:455-544 on CPU fp32.  ``prediction``: [B, rows, 5 + nc] (Detect's ``z``);
    returns one [n, 6] tensor (x1, y1, x2, y2, conf, cls) per image."""
    pred = prediction.detach().float().cpu().numpy()
    nc = pred.shape[2] - 5
    multi_label = bool(multi_label) and nc > 1                                            # :471
    ct = np.float32(conf_thres)
    out = []
    for x in pred:
        x = x[x[:, 4] > ct]                                                               # :458, :479
        if x.shape[0] == 0:
            out.append(torch.zeros((0, 6)))
            continue
        x = x.copy()
        x[:, 5:] *= x[:, 4:5]                                                             # :496
        box = xywh2xyxy(x[:, :4])                                                         # :499
        if multi_label:                                                                   # :502-504
            i, j = np.nonzero(x[:, 5:] > ct)
            x = np.concatenate([box[i], x[i, j + 5, None], j[:, None].astype(np.float32)], 1)
        else:                                                                             # :505-507
            j = x[:, 5:].argmax(1)
            conf = x[np.arange(x.shape[0]), 5 + j]
            x = np.concatenate([box, conf[:, None], j[:, None].astype(np.float32)], 1)[conf > ct]
        if classes is not None:                                                           # :510-511
            x = x[np.isin(x[:, 5], np.asarray(classes, np.float32))]
        n = x.shape[0]
        if n == 0:                                                                        # :519-520
            out.append(torch.zeros((0, 6)))
            continue
        if n > MAX_NMS:                                                                   # :521-522
            x = x[np.argsort(-x[:, 4], kind="stable")[:MAX_NMS]]
        c = x[:, 5:6] * np.float32(0 if agnostic else MAX_WH)                             # :525
        keep = nms_greedy(x[:, :4] + c, x[:, 4], iou_thres, limit=max_det)                # :526-529
        out.append(torch.from_numpy(x[keep]))
    return out


def make_predictions(b: int, rows: int, nc: int, seed: int, clusters: int = 40, img: float = 640.0,
                     conf_lo: float = 0.0) -> torch.Tensor:
    """Seeded synthetic Detect output [b, rows, 5 + nc] with heavily overlapping boxes (cluster centres + jitter), a
    spread of objectness around the 0.25 threshold and exact score ties (a few rows are duplicated), so that every
    branch of the suppression loop is exercised.  Same generator on the build container and on the GPU box."""
    g = torch.Generator().manual_seed(seed)
    ctr = torch.rand(b, clusters, 2, generator=g) * img
    size = 20 + torch.rand(b, clusters, 2, generator=g) * 120
    which = torch.randint(0, clusters, (b, rows), generator=g)
    idx = which.unsqueeze(-1).expand(-1, -1, 2)
    xy = torch.gather(ctr, 1, idx) + torch.randn(b, rows, 2, generator=g) * 6
    wh = torch.gather(size, 1, idx) * (0.8 + 0.4 * torch.rand(b, rows, 2, generator=g))
    obj = conf_lo + (1 - conf_lo) * torch.rand(b, rows, 1, generator=g)
    cls = torch.rand(b, rows, nc, generator=g)
    p = torch.cat([xy, wh, obj, cls], 2).float()
    ndup = max(1, rows // 50)                                  # exact duplicates -> score ties and IoU == 1
    src = torch.randint(0, rows, (ndup,), generator=g)
    dst = torch.randperm(rows, generator=g)[:ndup]             # unique targets: the scatter is order independent
    p[:, dst] = p[:, src]
    return p.contiguous()
