"""ctypes binding of ``libcft_b200.so`` (the C ABI declared in ``include/cft_b200.h``).

There is no CPU fallback: every op of the forward path goes through this library, and
``lib()`` raises if it is missing or the device is not sm_100.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcft_b200.so")
CSRC = os.path.join(_HERE, "csrc")

ACT_NONE, ACT_SILU, ACT_GELU = 0, 1, 2
DT_BF16, DT_F32, DT_U8 = 0, 1, 2
KERNEL_IDS = {
    "conv_tcgen05": 0, "conv_ref": 1, "focus": 2, "maxpool": 3, "upsample": 4, "add": 5, "copy": 6,
    "pool_tokens": 7, "layernorm": 8, "attention": 9, "unpool": 10, "detect": 11, "nms": 12, "gpt_block": 13,
}


class ConvArgs(C.Structure):
    """struct cft_conv_args (include/cft_b200.h)."""
    _fields_ = [
        ("x", C.c_void_p), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int),
        ("ldx", C.c_int), ("x_coff", C.c_int),
        ("w", C.c_void_p), ("bias", C.c_void_p), ("Cout", C.c_int), ("k", C.c_int), ("stride", C.c_int),
        ("act", C.c_int),
        ("res", C.c_void_p), ("ldr", C.c_int), ("r_coff", C.c_int),
        ("y", C.c_void_p), ("ldy", C.c_int), ("y_coff", C.c_int), ("out_dtype", C.c_int),
        ("kw", C.c_int),
        ("w2", C.c_void_p), ("bias2", C.c_void_p), ("y2", C.c_void_p), ("ldy2", C.c_int), ("y2_coff", C.c_int),
        ("act2", C.c_int), ("skip_y", C.c_int),
    ]


class ConvPlan(C.Structure):
    """struct cft_conv_plan (include/cft_b200.h): the launch plan cft_conv2d would use, computed on the host."""
    _fields_ = [(n, C.c_int) for n in (
        "ctas", "TW", "TH", "Ho", "Wo", "tiles_x", "tiles_y", "m_tiles", "block_n", "n_blocks", "num_tiles",
        "kelems", "kchunks", "ups", "halo", "stages", "a_slot", "b_slot", "b_res", "acc_stages", "acc_cols",
        "teams", "stage_c", "smem_bytes", "grid", "TB")]


class GptBlockArgs(C.Structure):
    """struct cft_gpt_block_args (include/cft_b200.h)."""
    _fields_ = [
        ("B", C.c_int), ("tokens", C.c_int), ("d", C.c_int), ("heads", C.c_int), ("layers", C.c_int),
        ("cluster", C.c_int),
        ("wqkv", C.c_void_p), ("bqkv", C.c_void_p), ("wo", C.c_void_p), ("bo", C.c_void_p),
        ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("ln1_g", C.c_void_p), ("ln1_b", C.c_void_p), ("ln2_g", C.c_void_p), ("ln2_b", C.c_void_p),
        ("lnf_g", C.c_void_p), ("lnf_b", C.c_void_p),
        ("eps1", C.c_float), ("eps2", C.c_float), ("epsf", C.c_float),
        ("x_in", C.c_void_p), ("x_out", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_longlong),
        ("debug_x", C.c_void_p),
    ]


_I, _P, _LL, _F = C.c_int, C.c_void_p, C.c_longlong, C.c_float
# name -> argtypes; every symbol include/cft_b200.h declares (tests check the export list).
SIGNATURES = {
    "cft_abi_version": ([], C.c_int),
    "cft_last_error": ([], C.c_char_p),
    "cft_check_device": ([C.POINTER(_I)] * 3, _I),
    "cft_conv2d": ([C.POINTER(ConvArgs), _P], _I),
    "cft_conv2d_ref": ([C.POINTER(ConvArgs), _P], _I),
    "cft_debug_conv_trace": ([_P], _I),
    "cft_debug_conv_spans": ([_P, _I], _I),
    "cft_debug_conv_plan": ([C.POINTER(ConvArgs), C.POINTER(ConvPlan)], _I),
    "cft_focus_gather": ([_P, _I, _I, _I, _I, _LL, _I, _P, _P], _I),
    "cft_focus_conv": ([_P, _I, _I, _I, _LL, _P, _P, _I, _I, _P, _I, _I, _P], _I),
    "cft_maxpool_s1": ([_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "cft_maxpool_cascade3": ([_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "cft_upsample2x": ([_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "cft_add": ([_P, _I, _I, _P, _I, _I, _P, _I, _I, _LL, _I, _P], _I),
    "cft_copy": ([_P, _I, _I, _P, _I, _I, _LL, _I, _P], _I),
    "cft_gpt_pool_tokens": ([_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P], _I),
    "cft_layernorm": ([_P, _P, _P, _F, _LL, _I, _P, _I, _P], _I),
    "cft_attention": ([_P, _P, _I, _I, _I, _I, _P], _I),
    "cft_gpt_block_workspace_bytes": ([_I, _I], _LL),
    "cft_gpt_block_supported": ([_I, _I, _I, _I], _I),
    "cft_gpt_block": ([C.POINTER(GptBlockArgs), _P], _I),
    "cft_debug_block_trace": ([_P], _I),
    "cft_gpt_unpool": ([_P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P, _I, _I, _P, _I, _I, _P, _I, _I, _P, _I, _I, _P], _I),
    "cft_detect_decode": ([_P, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P, _LL, _LL, _P], _I),
    "cft_nms_workspace_bytes": ([_I, _I, _I, _I], _LL),
    "cft_nms": ([_P, _I, _I, _I, _F, _F, _I, _I, _I, C.POINTER(_I), _I, _P, _LL, _P, _P, _P], _I),
    "cft_prof_enable": ([_I], _I),
    "cft_prof_get": ([_I, C.POINTER(C.c_double), C.POINTER(_LL)], _I),
    "cft_launch_count": ([], _LL),
}

_lib = None
_device_checked = False


class CftError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libcft_b200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise CftError("building libcft_b200.so failed")
    return LIB_PATH


def load(check_device: bool = False):
    """dlopen the library and attach prototypes. No device work unless check_device."""
    global _lib, _device_checked
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise CftError(f"{LIB_PATH} not found: run __graft_entry__.build() (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = lib
    if check_device and not _device_checked:
        sm, ma, mi = _I(), _I(), _I()
        rc = _lib.cft_check_device(C.byref(sm), C.byref(ma), C.byref(mi))
        if rc != 0:
            raise CftError("cft_check_device: " + _lib.cft_last_error().decode())
        _device_checked = True
    return _lib


def lib():
    return load(check_device=True)


def check(rc: int, what: str = ""):
    if rc != 0:
        raise CftError(f"{what} failed (code {rc}): {_lib.cft_last_error().decode()}")


def prof_enable(on: bool):
    check(lib().cft_prof_enable(1 if on else 0), "cft_prof_enable")


def prof_get():
    """{kernel name: (total_ms, launches)} since prof_enable(True)."""
    out = {}
    for name, kid in KERNEL_IDS.items():
        ms, n = C.c_double(), _LL()
        check(lib().cft_prof_get(kid, C.byref(ms), C.byref(n)), "cft_prof_get")
        out[name] = (ms.value, n.value)
    return out


def launch_count() -> int:
    return int(load().cft_launch_count())
