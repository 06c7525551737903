"""CPU suite: the reference arm of bench.py (`--impl reference`) -- the one bench leg that needs no GPU.  It must print exactly
one JSON line with the contract's keys, time the unmodified reference's own modules when the tree (or its staged copy) is
importable, and fall back to the oracle port when it is not."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[-1500:]
    return json.loads(lines[0])


@pytest.mark.parametrize("mode", ["default", "no_reference_tree"])
def test_reference_arm_prints_one_contract_line(mode):
    from oracle import ref_shim
    d = _run({} if mode == "default" else {"CFT_REFERENCE_ROOT": "/nonexistent"})
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("RGB+IR pairs/sec") and d["gpu_launches"] == 0 and d["dtype"] == "f32"
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["cores"] >= 1
    want = "reference" if (mode == "default" and ref_shim.available()) else "port"
    assert cb["kind"] == want, cb
