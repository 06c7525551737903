"""Model-graph descriptions for the two-stream ``*_fusion_transformerx3_*`` family.

The reference stores these graphs as yaml files under ``models/transformer/`` and
feeds them to ``parse_model`` (reference ``models/yolo_test.py:479-555``).  The yaml
files do not exist on the GPU box, so the same dictionaries are generated here;
``tests/test_host_cpu.py`` checks (when ``/root/reference`` is present) that
``x3_config(...)`` equals ``yaml.safe_load`` of every reference x3 yaml.

Row format is the reference's: ``[from, number, module-name, args]``.
``from == -4`` means "feed the IR image" (reference ``models/yolo_test.py:262-263``).
"""
from __future__ import annotations

import copy
import math

ANCHORS = [
    [10, 13, 16, 30, 33, 23],        # P3/8
    [30, 61, 62, 45, 59, 119],       # P4/16
    [116, 90, 156, 198, 373, 326],   # P5/32
]

# (depth_multiple, width_multiple); 'x' is derived: the reference
# ships no yolov5x x3 yaml, the multiples come from
# models/transformer/yolov5x_fusion_transformer_FLIR.yaml:3-4.
VARIANTS = {"s": (0.33, 0.50), "l": (1.0, 1.0), "x": (1.33, 1.25)}

# Named configs of BASELINE.json / SURVEY.md §8(d).
NAMED = {
    "yolov5s_fusion_transformerx3_vedai": ("s", 9),
    "yolov5l_fusion_transformerx3_FLIR_aligned": ("l", 3),
    "yolov5l_fusion_transformerx3_llvip": ("l", 1),
    "yolov5x_fusion_transformerx3_FLIR_aligned": ("x", 3),  # derived
}


def _stream(start_from):
    """One CSPDarknet stem up to P3 (yaml rows 0-4 / 5-9)."""
    return [
        [start_from, 1, "Focus", [64, 3]],
        [-1, 1, "Conv", [128, 3, 2]],
        [-1, 3, "C3", [128]],
        [-1, 1, "Conv", [256, 3, 2]],
        [-1, 9, "C3", [256]],
    ]


def x3_config(variant: str = "l", nc: int = 3) -> dict:
    """Return the model dict of ``yolov5{variant}_fusion_transformerx3`` with ``nc`` classes.

    Mirrors reference ``models/transformer/yolov5l_fusion_transformerx3_FLIR_aligned.yaml:1-97``.
    """
    gd, gw = VARIANTS[variant]
    backbone = []
    backbone += _stream(-1)                      # 0-4   RGB stream
    backbone += _stream(-4)                      # 5-9   IR stream
    backbone += [
        [[4, 9], 1, "GPT", [256]],               # 10 CFT @P3
        [[4, 10], 1, "Add2", [256, 0]],          # 11
        [[9, 10], 1, "Add2", [256, 1]],          # 12
        [11, 1, "Conv", [512, 3, 2]],            # 13
        [-1, 9, "C3", [512]],                    # 14
        [12, 1, "Conv", [512, 3, 2]],            # 15
        [-1, 9, "C3", [512]],                    # 16
        [[14, 16], 1, "GPT", [512]],             # 17 CFT @P4
        [[14, 17], 1, "Add2", [512, 0]],         # 18
        [[16, 17], 1, "Add2", [512, 1]],         # 19
        [18, 1, "Conv", [1024, 3, 2]],           # 20
        [-1, 1, "SPP", [1024, [5, 9, 13]]],      # 21
        [-1, 3, "C3", [1024, False]],            # 22
        [19, 1, "Conv", [1024, 3, 2]],           # 23
        [-1, 1, "SPP", [1024, [5, 9, 13]]],      # 24
        [-1, 3, "C3", [1024, False]],            # 25
        [[22, 25], 1, "GPT", [1024]],            # 26 CFT @P5
        [[22, 26], 1, "Add2", [1024, 0]],        # 27
        [[25, 26], 1, "Add2", [1024, 1]],        # 28
        [[11, 12], 1, "Add", [1]],               # 29 merged P3
        [[18, 19], 1, "Add", [1]],               # 30 merged P4
        [[27, 28], 1, "Add", [1]],               # 31 merged P5
    ]
    head = [
        [-1, 1, "Conv", [512, 1, 1]],                     # 32
        [-1, 1, "nn.Upsample", ["None", 2, "nearest"]],     # 33
        [[-1, 30], 1, "Concat", [1]],                     # 34
        [-1, 3, "C3", [512, False]],                      # 35
        [-1, 1, "Conv", [256, 1, 1]],                     # 36
        [-1, 1, "nn.Upsample", ["None", 2, "nearest"]],     # 37
        [[-1, 29], 1, "Concat", [1]],                     # 38
        [-1, 3, "C3", [256, False]],                      # 39  P3 out
        [-1, 1, "Conv", [256, 3, 2]],                     # 40
        [[-1, 36], 1, "Concat", [1]],                     # 41
        [-1, 3, "C3", [512, False]],                      # 42  P4 out
        [-1, 1, "Conv", [512, 3, 2]],                     # 43
        [[-1, 32], 1, "Concat", [1]],                     # 44
        [-1, 3, "C3", [1024, False]],                     # 45  P5 out
        [[39, 42, 45], 1, "Detect", ["nc", "anchors"]],   # 46
    ]
    return {
        "nc": nc,
        "depth_multiple": gd,
        "width_multiple": gw,
        "anchors": copy.deepcopy(ANCHORS),
        "backbone": backbone,
        "head": head,
    }


def named_config(name: str) -> dict:
    variant, nc = NAMED[name]
    return x3_config(variant, nc)


def make_divisible(x, divisor):
    """reference utils/general.py:210-212"""
    return math.ceil(x / divisor) * divisor
