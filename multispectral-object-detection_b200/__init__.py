"""B200-native (sm_100a) two-stream CFT / yolov5-CFTx3 forward path.

Drop-in for ONE hot path of DocF/multispectral-object-detection: the
``yolov5{s,l,x}_fusion_transformerx3`` forward (dual CSPDarknet backbones -> 3 CFT/GPT fusion
blocks -> PANet neck -> Detect).  Hand-written CUDA behind a C ABI (``include/cft_b200.h``,
``libcft_b200.so``), ``nn.Module`` mirrors of the reference classes on top.

The directory name carries a hyphen (it mirrors the reference repository's name); import it with
``importlib.import_module("multispectral-object-detection_b200")`` or through the root alias
module ``cft_b200``.
"""
from . import _lib, allreduce, checkpoint, config, nms, ops, shard  # noqa: F401
from .checkpoint import attempt_load, from_reference_model  # noqa: F401
from ._lib import CftError, build, load  # noqa: F401
from .config import named_config, x3_config  # noqa: F401
from .engine import ForwardEngine  # noqa: F401
from .model import Model, convert, install, parse_model, uninstall  # noqa: F401
from .nms import nms_batched, non_max_suppression  # noqa: F401
from .modules import (C3, SPP, Add, Add2, Bottleneck, Concat, Conv, Detect, Focus, GPT,  # noqa: F401
                      Upsample)

__all__ = ["Model", "ForwardEngine", "install", "uninstall", "convert", "parse_model", "x3_config", "named_config",
           "Conv", "Focus", "Bottleneck", "C3", "SPP", "Concat", "Add", "Add2", "GPT", "Detect", "Upsample",
           "build", "load", "CftError", "ops", "config", "nms", "non_max_suppression", "nms_batched", "shard", "allreduce", "checkpoint", "attempt_load", "from_reference_model"]
