"""CPU suite for the NMS row: the CPU restatement ``oracle/nms_oracle.py`` against the committed
golden outputs of the UNMODIFIED reference ``utils.general.non_max_suppression`` (oracle/make_golden_nms.py) and, when
/root/reference is present, against the reference function itself on fresh seeds.  Bit-exact."""
import os

import pytest
import torch

from oracle import nms_oracle as N
from oracle import ref_shim


def _golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "nms_cases.pt"))


def test_nms_golden_covers_every_branch(golden_dir):
    g = _golden(golden_dir)
    kws = [c["nms_kwargs"] for c in g.values()]
    assert any(k.get("multi_label") for k in kws) and any(k.get("agnostic") for k in kws)
    assert any(k.get("classes") for k in kws)
    assert any(sum(o.shape[0] for o in c["out"]) == 0 for c in g.values())                 # nothing passes
    assert any(c["args"][1] * (c["args"][2] if c["nms_kwargs"].get("multi_label") else 1) > N.MAX_NMS
               for c in g.values())                                                         # > max_nms candidates
    assert any(any(o.shape[0] == N.MAX_DET for o in c["out"]) for c in g.values())         # max_det cap hit


def test_nms_oracle_matches_reference_golden(golden_dir):
    for name, c in _golden(golden_dir).items():
        b, rows, nc, seed = c["args"]
        p = N.make_predictions(b, rows, nc, seed, **c["pred_kwargs"])
        assert abs(float(p.double().sum()) - c["input_checksum"]) < 1e-6, name     # same inputs as the reference saw
        out = N.non_max_suppression(p, **c["nms_kwargs"])
        assert len(out) == len(c["out"])
        for a, r in zip(out, c["out"]):
            assert a.shape == r.shape and torch.equal(a, r), name


def test_nms_oracle_properties():
    p = N.make_predictions(2, 4000, 4, seed=11)
    out = N.non_max_suppression(p, conf_thres=0.3, iou_thres=0.5)
    for d in out:
        assert d.shape[0] <= N.MAX_DET and (d[:, 4] > 0.3).all()
        assert (d[1:, 4] <= d[:-1, 4]).all()                                        # descending confidence
        # no kept pair of one class overlaps by more than the threshold
        import torchvision
        iou = torchvision.ops.box_iou(d[:, :4], d[:, :4])
        same = d[:, 5:6] == d[:, 5:6].T
        iou = iou * same - torch.eye(d.shape[0])
        assert iou.max() <= 0.5 + 1e-6


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("kw", [{}, {"multi_label": True}, {"agnostic": True, "iou_thres": 0.3}, {"classes": [0]}])
def test_nms_oracle_equals_live_reference(kw):
    ref_shim.import_reference()
    from utils.general import non_max_suppression as ref_nms
    p = N.make_predictions(2, 2000, 3, seed=21)
    ref = ref_nms(p.clone(), **kw)
    out = N.non_max_suppression(p, **kw)
    for a, r in zip(out, ref):
        assert a.shape == r.shape and torch.equal(a, r)


def test_nms_wrapper_fails_loudly_without_cuda(cft):
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from importlib import import_module
    nms = import_module("multispectral-object-detection_b200.nms")
    with pytest.raises(cft.CftError):
        nms.non_max_suppression(torch.rand(1, 10, 8))
