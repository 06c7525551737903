"""``non_max_suppression`` of the reference (``utils/general.py:455-544``) on the B200: one kernel launch for the
whole batch (``cft_nms``: confidence filter, class selection, xywh->xyxy, per-image stable sort, greedy IoU
suppression with per-class box offsets, max_det cap), bit-identical to the reference + ``torchvision.ops.nms``.

Same signature, defaults and return value as the reference function (a list with one ``[n, 6]`` tensor
``x1, y1, x2, y2, conf, cls`` per image), so ``detect_twostream.py:86`` / ``test.py:129`` can call it unchanged.
``nms_batched`` is the graph-capturable form: fixed-shape ``[B, max_det, 6]`` + ``counts [B]``, no host sync.
CUDA only -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import CftError

MAX_DET = 300          # utils/general.py:465


def nms_batched(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45,
                classes: Optional[Sequence[int]] = None, agnostic: bool = False, multi_label: bool = False,
                max_det: int = MAX_DET, out: Optional[torch.Tensor] = None, counts: Optional[torch.Tensor] = None,
                workspace: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """``prediction`` fp32 [B, rows, 5 + nc] on the GPU -> (det fp32 [B, max_det, 6], counts int32 [B]);
    rows ``[0, counts[b])`` of image b are valid, in descending confidence.  No synchronisation."""
    lib = _lib.lib()
    if not prediction.is_cuda:
        raise CftError(f"nms: prediction is on {prediction.device}; the CFT path is CUDA (sm_100a) only -- no CPU fallback")
    if prediction.dim() != 3 or prediction.shape[2] <= 5:
        raise CftError(f"nms: expected [B, rows, 5 + nc], got {tuple(prediction.shape)}")
    if prediction.dtype != torch.float32 or not prediction.is_contiguous():
        prediction = prediction.float().contiguous()
    b, rows, no = prediction.shape
    dev = prediction.device
    if b == 0 or rows == 0:
        return (torch.zeros((b, max_det, 6), dtype=torch.float32, device=dev),
                torch.zeros((b,), dtype=torch.int32, device=dev))
    ws_bytes = int(lib.cft_nms_workspace_bytes(b, rows, no - 5, 1 if multi_label else 0))
    if workspace is None or workspace.numel() * workspace.element_size() < ws_bytes:
        workspace = torch.empty(((ws_bytes + 7) // 8,), dtype=torch.int64, device=dev)
    if out is None:
        out = torch.zeros((b, max_det, 6), dtype=torch.float32, device=dev)
    if counts is None:
        counts = torch.zeros((b,), dtype=torch.int32, device=dev)
    if tuple(out.shape) != (b, max_det, 6) or out.dtype != torch.float32 or not out.is_contiguous():
        raise CftError("nms: out must be contiguous fp32 [B, max_det, 6]")
    cls_list = [int(c) for c in classes] if classes is not None else []
    if classes is not None and not cls_list:
        cls_list = [-1]                                   # empty filter keeps nothing (as the reference's isin)
    arr = (C.c_int * max(1, len(cls_list)))(*cls_list) if cls_list else None
    _lib.check(lib.cft_nms(prediction.data_ptr(), b, rows, no, float(conf_thres), float(iou_thres), int(max_det),
                           1 if multi_label else 0, 1 if agnostic else 0, arr, len(cls_list),
                           workspace.data_ptr(), workspace.numel() * workspace.element_size(),
                           out.data_ptr(), counts.data_ptr(), torch.cuda.current_stream().cuda_stream), "cft_nms")
    return out, counts


def non_max_suppression(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45,
                        classes: Optional[Sequence[int]] = None, agnostic: bool = False, multi_label: bool = False,
                        labels=()) -> List[torch.Tensor]:
    """Drop-in for reference ``utils/general.py:455`` (same arguments, same list-of-[n,6] result).  One device->host
    read of the per-image counts is the only synchronisation (the reference's result shapes are data dependent)."""
    if labels is not None and len(labels):
        raise CftError("non_max_suppression: the `labels` (autolabelling) branch (utils/general.py:482-489) is a "
                       "training-time feature outside the inference hot path")
    det, counts = nms_batched(prediction, conf_thres, iou_thres, classes, agnostic, multi_label)
    n = counts.cpu().tolist()
    return [det[i, :k] for i, k in enumerate(n)]
