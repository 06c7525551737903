"""CPU suite: checkpoint ingestion.  A checkpoint written exactly as the reference's
``train.py:850-857`` writes it (the UNMODIFIED reference ``Model``, ``.half()``, pickled whole) must load through
``attempt_load`` (mirror of ``models/experimental.py:113-134``) into a B200 ``Model`` with identical weights -- both with the
reference tree importable and, in a fresh interpreter WITHOUT it, through the ``models.*`` alias modules."""
import json
import os
import subprocess
import sys

import pytest
import torch

from oracle import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAME = "yolov5s_fusion_transformerx3_vedai"


@pytest.fixture(scope="module")
def checkpoint(tmp_path_factory, cft, oracle):
    if not ref_shim.available():
        pytest.skip("/root/reference not present (GPU box): nothing can write a reference checkpoint")
    yt = ref_shim.import_reference()
    cfg = cft.named_config(NAME)
    model = yt.Model(ref_shim.reference_yaml(NAME), ch=3)
    sd = oracle.init_state(cfg, seed=7)
    model.load_state_dict(sd, strict=True)
    model.names = [f"cls{i}" for i in range(cfg["nc"])]
    path = str(tmp_path_factory.mktemp("ckpt") / "last.pt")
    from copy import deepcopy
    ckpt = {"epoch": 3, "best_fitness": 0.5, "training_results": "", "model": deepcopy(model).half(), "ema": None,
            "updates": 0, "optimizer": None, "wandb_id": None}                       # train.py:850-857
    torch.save(ckpt, path)
    return path, {k: v.half().float() if v.is_floating_point() else v for k, v in sd.items()}


def _summary(model):
    sd = model.state_dict()
    return {"layers": [type(m).__name__ for m in model.model], "modules": [type(m).__module__ for m in model.model],
            "n_keys": len(sd), "checksum": float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point())),
            "names": model.names, "stride": [float(s) for s in model.stride], "save": list(model.save),
            "gpt_groups": {str(k): v for k, v in model._plan["gpt_groups"].items()}}


def test_attempt_load_with_reference_importable(checkpoint, cft):
    path, sd_half = checkpoint
    model = cft.attempt_load(path, fuse=False)
    assert isinstance(model, cft.Model) and not model.training
    got = model.state_dict()
    assert set(got) == set(sd_half)
    for k, v in sd_half.items():
        assert torch.equal(got[k].float(), v.float()), k
    assert all(type(m).__module__.startswith("multispectral-object-detection_b200") or type(m).__name__ == "Upsample"
               for m in model.model)
    assert model._plan["gpt_groups"] and model.names[0] == "cls0"
    fused = cft.attempt_load(path)                                         # default: .fuse() like the reference
    assert not any(hasattr(m, "bn") for m in fused.modules() if type(m).__name__ == "Conv")


def test_attempt_load_without_the_reference_tree(checkpoint, cft):
    path, _ = checkpoint
    ref = _summary(cft.attempt_load(path, fuse=False))
    code = (
        "import sys, json, importlib\n"
        f"sys.path = [p for p in sys.path if 'reference' not in p]; sys.path.insert(0, {ROOT!r})\n"
        "assert 'models' not in sys.modules\n"
        "cft = importlib.import_module('multispectral-object-detection_b200')\n"
        "sys.path.insert(0, %r)\n" % os.path.join(ROOT, "tests") +
        "from test_checkpoint_cpu import _summary\n"
        f"m = cft.attempt_load({path!r}, fuse=False)\n"
        "assert 'models' not in sys.modules, 'alias modules must not leak'\n"
        "print('SUMMARY ' + json.dumps(_summary(m)))\n")
    env = {k: v for k, v in os.environ.items() if k not in ("PYTHONPATH", "CFT_REFERENCE_ROOT")}
    env["CFT_REFERENCE_ROOT"] = "/nonexistent"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600, cwd="/tmp")
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("SUMMARY ")][-1]
    got = json.loads(line[len("SUMMARY "):])
    assert got == json.loads(json.dumps(ref))


def test_attempt_load_rejects_ensembles(cft):
    with pytest.raises(cft.CftError):
        cft.attempt_load(["a.pt", "b.pt"])
