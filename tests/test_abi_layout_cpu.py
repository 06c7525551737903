"""CPU suite: the ctypes mirrors in `_lib.py` have exactly the layout of the C structs in include/cft_b200.h.  A C probe
compiled with gcc from the header prints sizeof / offsetof of every field; the ctypes classes must agree (an appended or
re-ordered field on one side only would otherwise corrupt arguments silently)."""
import ctypes as C
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STRUCTS = {"cft_conv_args": "ConvArgs", "cft_conv_plan": "ConvPlan", "cft_gpt_block_args": "GptBlockArgs"}


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_ctypes_structs_match_the_header(cft, tmp_path):
    L = cft._lib
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "cft_b200.h"', "int main(void) {"]
    for cname, pyname in STRUCTS.items():
        cls = getattr(L, pyname)
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "probe"
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr          # also: every mirrored field exists in the header under the same name
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    seen = {}
    for ln in out.splitlines():
        cname, field, val = ln.split()
        seen[(cname, field)] = int(val)
    for cname, pyname in STRUCTS.items():
        cls = getattr(L, pyname)
        assert seen[(cname, "size")] == C.sizeof(cls), (cname, seen[(cname, "size")], C.sizeof(cls))
        for fname, _ in cls._fields_:
            assert seen[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
    # and the header declares no field the mirror lacks: the sizes agree and the last mirrored field ends at the struct's end
    for cname, pyname in STRUCTS.items():
        cls = getattr(L, pyname)
        last, ltype = cls._fields_[-1]
        assert getattr(cls, last).offset + C.sizeof(ltype) <= C.sizeof(cls) < getattr(cls, last).offset + C.sizeof(ltype) + 8


def _kind(ctype):
    """int / ll / float / ptr for a ctypes argtype."""
    if ctype is C.c_int:
        return "int"
    if ctype is C.c_longlong:
        return "ll"
    if ctype is C.c_float:
        return "float"
    return "ptr"


def test_ctypes_signatures_match_the_header_prototypes(cft):
    """Every prototype of include/cft_b200.h, parsed: parameter count and kind (int / long long / float / pointer) and the
    return type must equal the ctypes table the Python side calls through (`_lib.SIGNATURES`)."""
    import re
    L = cft._lib
    text = open(os.path.join(ROOT, "include", "cft_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = re.findall(r"^\s*(int|long long|const char\*)\s+(cft_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.M | re.S)
    assert len(protos) >= 25
    names = set()
    for ret, name, params in protos:
        names.add(name)
        assert name in L.SIGNATURES, f"{name} is declared in the header but has no ctypes signature"
        argtypes, restype = L.SIGNATURES[name]
        params = " ".join(params.split())
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        kinds = []
        for p in plist:
            if "*" in p:
                kinds.append("ptr")
            elif p.startswith("long long"):
                kinds.append("ll")
            elif p.startswith("float"):
                kinds.append("float")
            else:
                assert p.startswith("int "), (name, p)
                kinds.append("int")
        assert kinds == [_kind(t) for t in argtypes], (name, kinds, [_kind(t) for t in argtypes])
        want_ret = {"int": C.c_int, "long long": C.c_longlong, "const char*": C.c_char_p}[ret]
        assert restype is want_ret, (name, ret, restype)
    assert names == set(L.SIGNATURES), sorted(set(L.SIGNATURES) ^ names)


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_enum_constants_match_the_header(cft, tmp_path):
    L = cft._lib
    kid = {"conv_tcgen05": "CFT_K_CONV_TCGEN05", "conv_ref": "CFT_K_CONV_REF", "focus": "CFT_K_FOCUS", "maxpool": "CFT_K_MAXPOOL",
           "upsample": "CFT_K_UPSAMPLE", "add": "CFT_K_ADD", "copy": "CFT_K_COPY", "pool_tokens": "CFT_K_POOL_TOKENS",
           "layernorm": "CFT_K_LAYERNORM", "attention": "CFT_K_ATTENTION", "unpool": "CFT_K_UNPOOL", "detect": "CFT_K_DETECT",
           "nms": "CFT_K_NMS", "gpt_block": "CFT_K_GPT_BLOCK"}
    assert set(kid) == set(L.KERNEL_IDS)
    consts = {"CFT_ACT_NONE": L.ACT_NONE, "CFT_ACT_SILU": L.ACT_SILU, "CFT_ACT_GELU": L.ACT_GELU, "CFT_DT_BF16": L.DT_BF16,
              "CFT_DT_F32": L.DT_F32, "CFT_DT_U8": L.DT_U8, "CFT_K_COUNT": len(L.KERNEL_IDS)}
    consts.update({c: L.KERNEL_IDS[k] for k, c in kid.items()})
    body = "\n".join(f'  printf("{c} %d\\n", (int){c});' for c in list(consts) + ["CFT_ABI_VERSION"])
    src = tmp_path / "enums.c"
    src.write_text('#include <stdio.h>\n#include "cft_b200.h"\nint main(void) {\n' + body + "\n  return 0;\n}\n")
    exe = tmp_path / "enums"
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for c, v in consts.items():
        assert int(got[c]) == v, (c, got[c], v)
    assert int(got["CFT_ABI_VERSION"]) == cft.load().cft_abi_version()        # the built library is of this header
