"""Generate tests/golden/*.pt from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE.  For each case: build the reference ``models.yolo_test.Model`` from the
reference's own yaml (or, for the derived yolov5x x3 graph, from the dict), load the seeded
synthetic state (``cft_oracle.init_state``), run the reference eval forward on CPU fp32,
assert the CPU restatement (``cft_oracle.forward``) reproduces it, and store the REFERENCE's
outputs plus float64 checksums of the seeded inputs/weights so the GPU box (which has no
``/root/reference``) can verify it regenerated the same tensors.

    python oracle/make_golden.py          # writes tests/golden/*.pt
"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cft_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

config = importlib.import_module("multispectral-object-detection_b200.config")

# name, config-name, batch, H, W, weight seed, input seed, fused
CASES = [
    ("s_vedai_b2_128x160", "yolov5s_fusion_transformerx3_vedai", 2, 128, 160, 0, 1, False),
    ("s_vedai_b1_64x64_fused", "yolov5s_fusion_transformerx3_vedai", 1, 64, 64, 3, 4, True),
    ("l_flir_b1_64x64", "yolov5l_fusion_transformerx3_FLIR_aligned", 1, 64, 64, 0, 1, False),
    ("l_llvip_b1_64x96", "yolov5l_fusion_transformerx3_llvip", 1, 64, 96, 5, 6, False),
    ("x_flir_b1_64x64", "yolov5x_fusion_transformerx3_FLIR_aligned", 1, 64, 64, 0, 1, False),
]


def checksum(t):
    return float(t.double().sum())


def state_checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def main():
    yt = ref_shim.import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, cname, b, h, w, wseed, iseed, fused in CASES:
        cfg = config.named_config(cname)
        ypath = ref_shim.reference_yaml(cname)
        model = yt.Model(ypath if os.path.isfile(ypath) else cfg, ch=3).eval()
        sd = O.init_state(cfg, seed=wseed)
        model.load_state_dict(sd, strict=True)
        if fused:
            model.fuse()                      # models/yolo_test.py:296-304
        x, x2 = O.make_inputs(b, h, w, seed=iseed)
        with torch.no_grad():
            z_ref, raw_ref = model(x, x2)
        osd = {k: v for k, v in model.state_dict().items()} if fused else sd
        z_o, raw_o = O.forward(osd, cfg, x, x2)
        dz = (z_ref - z_o).abs().max().item()
        dr = max((a - c).abs().max().item() for a, c in zip(raw_ref, raw_o))
        assert dz <= 1e-4 and dr <= 1e-5, (name, dz, dr)
        torch.save({
            "case": name, "config": cname, "batch": b, "height": h, "width": w,
            "weight_seed": wseed, "input_seed": iseed, "fused": fused,
            "z": z_ref.clone(), "raw": [r.clone() for r in raw_ref],
            "input_checksum": [checksum(x), checksum(x2)],
            "state_checksum": state_checksum(sd),
            "oracle_vs_reference_max_abs": [dz, dr],
            "torch": str(torch.__version__),
        }, os.path.join(out_dir, name + ".pt"))
        print(f"{name}: z{tuple(z_ref.shape)} oracle-vs-reference max|d| z={dz:.2e} raw={dr:.2e}")


if __name__ == "__main__":
    main()
