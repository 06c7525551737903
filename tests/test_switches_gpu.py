"""The library's environment switches are read once at load; every NON-default arm is exercised here in a child process
against the same parity tests as the default arm ("test the non-default numerics switches or delete them" -- the
pure timing-experiment switches of round 1 were deleted).  Tolerances are those of the tests that are re-run."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = os.path.join("tests", "test_kernels_gpu.py")
MODEL = os.path.join("tests", "test_model_gpu.py")

ARMS = [
    # environment, test file, -k filter
    ({"CFT_SILU_EXP2": "1"}, KERNELS, "conv_tcgen05 or chained"),          # SiLU as x * rcp(1 + 2^-x) instead of h + h tanh(h)
    ({"CFT_GELU_ERFF": "1"}, KERNELS, "gemm_linear"),                      # erff instead of the A&S erf-GELU
    ({"CFT_NO_ROW_REUSE": "1"}, KERNELS, "conv_tcgen05 or chained"),       # every 3x3 on the plain one-tap-per-stage path
    ({"CFT_NO_BRES": "1"}, KERNELS, "conv_tcgen05 or chained"),            # no resident 3x3 weights
    ({"CFT_NO_PDL": "1"}, KERNELS, "conv_tcgen05 or gemm_linear"),         # no programmatic dependent launch
    ({"CFT_CONV_CTAS": "1"}, KERNELS, "conv_tcgen05 or chained or gemm_linear"),   # never CTA pairs
    ({"CFT_CONV_CTAS": "2"}, KERNELS, "conv_tcgen05 or chained or gemm_linear"),   # CTA pairs wherever legal
    ({"CFT_ATTENTION_SIMT": "1"}, KERNELS, "attention_core"),              # the CUDA-core attention cross-check kernel
    ({"CFT_NO_BATCH_TILES": "1"}, KERNELS, "conv_tcgen05 or chained"),     # tiles never span images (the round-1 tiling)
    ({"CFT_NO_CONV_CHAIN": "1"}, MODEL, "golden"),                         # every Bottleneck 1x1 launched separately
    ({"CFT_NO_FUSED_BLOCK": "1"}, MODEL, "golden"),                        # CFT blocks on the per-op path
    ({"CFT_FUSED_BLOCK_MAX_D": "512"}, MODEL, "golden"),                   # ... and the one-launch kernel up to d = 512
    ({"CFT_ONE_STREAM": "1", "CFT_NO_FUSED_FOCUS": "1"}, MODEL, "golden or engine"),   # one stream; gather + conv Focus
]


@pytest.mark.parametrize("env,path,expr", ARMS, ids=["+".join(f"{k}={v}" for k, v in a[0].items()) for a in ARMS])
def test_non_default_arm(env, path, expr):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-m", "pytest", path, "-x", "-q", "-m", "gpu", "-k", expr], cwd=ROOT, env=e,
                       capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
