"""GPU parity of ``cft_nms`` (through the C ABI) -- bit-exact index/byte work:
(a) the committed golden outputs of the UNMODIFIED reference function, (b) the CPU oracle on fresh seeds and on the
sizes the forward produces (25200 rows @640x640, 80640 rows @1024x1280), (c) size-independent properties at the
full batch (sortedness, no surviving overlap, idempotence, batch invariance), (d) CUDA-graph capture, error paths."""
import os
from importlib import import_module

import pytest
import torch

from oracle import nms_oracle as N

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def nms():
    return import_module("multispectral-object-detection_b200.nms")


def _same(out, ref, what):
    assert len(out) == len(ref), what
    for i, (a, r) in enumerate(zip(out, ref)):
        a = a.cpu()
        assert a.shape == r.shape, f"{what}: image {i} kept {a.shape[0]} vs {r.shape[0]}"
        assert torch.equal(a, r), f"{what}: image {i} max|d| {float((a - r).abs().max())}"


def test_nms_matches_reference_golden(nms, golden_dir):
    g = torch.load(os.path.join(golden_dir, "nms_cases.pt"))
    for name, c in g.items():
        b, rows, nc, seed = c["args"]
        p = N.make_predictions(b, rows, nc, seed, **c["pred_kwargs"])
        assert abs(float(p.double().sum()) - c["input_checksum"]) < 1e-6, name
        _same(nms.non_max_suppression(p.to(DEV), **c["nms_kwargs"]), c["out"], name)


@pytest.mark.parametrize("b,rows,nc,seed,kw", [
    (4, 25200, 3, 31, {}),                                              # 640x640, FLIR
    (2, 25200, 9, 32, {"multi_label": True, "conf_thres": 0.1}),        # vedai, > 30000 candidates -> workspace sort
    (2, 80640, 1, 33, {"conf_thres": 0.001}),                           # LLVIP 1024x1280: > smem, > max_nms
    (3, 1000, 3, 34, {"iou_thres": 0.0}),
    (3, 1000, 3, 35, {"iou_thres": 1.0}),
    (1, 33, 2, 36, {"conf_thres": 0.0}),
    (5, 7, 80, 37, {"multi_label": True, "conf_thres": 0.01}),          # COCO-sized class count
])
def test_nms_matches_oracle(nms, b, rows, nc, seed, kw):
    p = N.make_predictions(b, rows, nc, seed)
    _same(nms.non_max_suppression(p.to(DEV), **kw), N.non_max_suppression(p, **kw), f"{b}x{rows}x{nc} {kw}")


def test_nms_exact_ties_and_duplicates(nms):
    """All scores equal, many identical boxes: order must be the stable (row-index) order."""
    p = N.make_predictions(2, 512, 2, seed=40)
    p[..., 4] = 0.75
    p[..., 5:] = 0.5
    p[:, 100:200, :4] = p[:, 300:400, :4]
    _same(nms.non_max_suppression(p.to(DEV)), N.non_max_suppression(p), "ties")


def test_nms_full_batch_properties(nms):
    """Batch 32 x 25200 rows (the bench configuration's z): properties that need no oracle."""
    p = N.make_predictions(32, 25200, 3, seed=41).to(DEV)
    det, counts = nms.nms_batched(p)
    torch.cuda.synchronize()
    counts = counts.cpu()
    assert (counts > 0).all() and (counts <= 300).all()
    import torchvision
    for i in range(0, 32, 5):
        d = det[i, :counts[i]].cpu()
        assert (d[:, 4] > 0.25).all() and (d[1:, 4] <= d[:-1, 4]).all()
        iou = torchvision.ops.box_iou(d[:, :4], d[:, :4]) * (d[:, 5:6] == d[:, 5:6].T) - torch.eye(d.shape[0])
        assert iou.max() <= 0.45 + 1e-5
    # batch invariance: image 7 alone gives the same rows
    d7, c7 = nms.nms_batched(p[7:8].contiguous())
    assert int(c7[0]) == int(counts[7]) and torch.equal(d7[0, :int(c7[0])], det[7, :int(counts[7])])
    # idempotence: the kept boxes survive a second pass unchanged (xyxy -> xywh is exact for these values or off by
    # one rounding, so compare the kept SET size and scores)
    k = int(counts[3])
    d = det[3, :k]
    again = torch.zeros(1, k, 8, device=DEV)
    again[0, :, 0] = (d[:, 0] + d[:, 2]) / 2
    again[0, :, 1] = (d[:, 1] + d[:, 3]) / 2
    again[0, :, 2] = d[:, 2] - d[:, 0]
    again[0, :, 3] = d[:, 3] - d[:, 1]
    again[0, :, 4] = 1.0
    again[0, torch.arange(k), 5 + d[:, 5].long()] = d[:, 4]
    d2, c2 = nms.nms_batched(again, iou_thres=0.46)
    assert int(c2[0]) == k and torch.equal(d2[0, :k, 4], d[:, 4])


def test_nms_graph_capture_and_reuse(nms):
    p = N.make_predictions(2, 3000, 3, seed=42).to(DEV)
    out = torch.zeros(2, 300, 6, device=DEV)
    counts = torch.zeros(2, dtype=torch.int32, device=DEV)
    ws = torch.empty(2 * 3000, dtype=torch.int64, device=DEV)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        nms.nms_batched(p, out=out, counts=counts, workspace=ws)       # warm-up
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            nms.nms_batched(p, out=out, counts=counts, workspace=ws)
        p.copy_(N.make_predictions(2, 3000, 3, seed=43).to(DEV))
        g.replay()
    s.synchronize()
    ref = N.non_max_suppression(N.make_predictions(2, 3000, 3, seed=43))
    _same([out[i, :int(counts[i])] for i in range(2)], ref, "graph replay")


def test_nms_errors(nms, cft):
    with pytest.raises(cft.CftError):
        nms.non_max_suppression(torch.rand(1, 10, 8))                   # CPU tensor: no fallback
    with pytest.raises(cft.CftError):
        nms.non_max_suppression(torch.rand(1, 10, 8, device=DEV), labels=[torch.zeros(1, 5)])
    with pytest.raises(cft.CftError):
        nms.nms_batched(torch.rand(1, 10, 8, device=DEV), max_det=5000)
    out = nms.non_max_suppression(torch.rand(2, 10, 8, device=DEV), classes=[])
    assert all(o.shape == (0, 6) for o in out)


def test_forward_engine_with_nms_stage(nms, cft):
    """ForwardEngine(nms=...) captures forward + NMS in one graph: its (det, counts) equal the oracle NMS of the z the
    plain engine returns for the same batch (detect_twostream.py:83-86 order of operations)."""
    from oracle import cft_oracle as O
    cfg = cft.named_config("yolov5s_fusion_transformerx3_vedai")
    model = cft.Model(cfg).eval()
    model.load_state_dict(O.init_state(cfg, seed=13), strict=True)
    model = model.to(DEV)
    g = torch.Generator().manual_seed(5)
    batches = [torch.randint(0, 256, (2, 6, 96, 128), dtype=torch.uint8, generator=g).pin_memory() for _ in range(3)]
    kw = {"conf_thres": 0.05, "iou_thres": 0.45}
    plain = cft.ForwardEngine(model, 2, 96, 128, device=DEV)
    fused = cft.ForwardEngine(model, 2, 96, 128, device=DEV, nms=kw)
    assert fused.launches_per_forward == plain.launches_per_forward + 1
    for hb in batches:
        z = plain.infer(hb).clone()
        det, counts = fused.infer(hb)
        ref = N.non_max_suppression(z, **kw)
        assert sum(r.shape[0] for r in ref) > 0
        _same([det[i, :int(counts[i])] for i in range(2)], ref, "engine+nms")
