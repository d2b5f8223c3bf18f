"""C-ABI library: loads, exports every symbol include/snacb.h declares, host-only entry points work.
No compute calls here (no GPU in the authoring container)."""
import ctypes as C
import os
import re

import pytest

from project_morpheus_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "snacb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(snacb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(ensure_lib):
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/snacb.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"


def test_header_constants_match_the_ctypes_binding():
    """Every SNACB_FLAG_* / status / precision / noise constant of include/snacb.h has the same value in _lib.py."""
    src = open(os.path.join(ROOT, "include", "snacb.h")).read()
    defs = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"^#define\s+SNACB_(\w+)\s+(-?(?:0x)?[0-9a-fA-F]+)\b", src, flags=re.M)}
    assert len(defs) >= 15
    checked = 0
    for name, val in defs.items():
        if hasattr(_lib, name):
            assert getattr(_lib, name) == val, name
            checked += 1
    for flag in ("FLAG_NO_RU_FUSION", "FLAG_NO_CONVT_NOISE_FUSION", "FLAG_PERSISTENT_RU", "FLAG_TAIL_FUSION",
                 "FLAG_NO_PERSISTENT_CONVT", "FLAG_FUSE_RU256", "FLAG_NO_BLOCK_FUSION", "FLAG_NO_CONVT_NOISE_COMPOSE"):
        assert flag in defs and getattr(_lib, flag) == defs[flag]
    assert checked >= 10
    assert _lib.PREC_FP16X3 == defs["PREC_FP16X3"] and _lib.WIN_NONFINITE == defs["WIN_NONFINITE"]


def test_plan_matches_dependency_cone(ensure_lib):
    """rows needed for samples [2048,4096) of a 4-frame and a 7-frame window."""
    lib = _lib.load()
    r = (C.c_int32 * 52)()
    for frames in (4, 7):
        assert lib.snacb_plan(frames, 2048, 4096, 1, r) == 0
        v = list(r)
        pairs = [tuple(v[i:i + 2]) for i in range(0, 52, 2)]
        z, h = pairs[0], pairs[1]
        blocks = [pairs[2 + 6 * b: 8 + 6 * b] for b in range(4)]
        assert h == (0, 15)                      # block-0 input
        assert blocks[0][2] == (0, 111)          # block-0 ConvT output
        assert blocks[1][0] == (24, 72) and blocks[1][2] == (201, 567)
        assert blocks[2][0] == (240, 528) and blocks[2][2] == (963, 2109)
        assert blocks[3][0] == (1002, 2070) and blocks[3][2] == (2006, 4138)
        assert blocks[3][5] == (2045, 4099)      # block-3 output feeding the k7 tail
        assert z == ((0, 16) if frames == 4 else (0, 18))
    assert lib.snacb_plan(0, 0, 1, 1, r) == _lib.OK - 1  # SNACB_EINVAL


def test_create_fails_loudly_without_gpu(ensure_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = C.c_void_p()
    cfg = _lib.Config(abi_version=_lib.ABI_VERSION, device=0, precision=_lib.PREC_FP32, chunk_items=0, trim=1)
    rc = lib.snacb_create(C.byref(h), C.byref(cfg))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in lib.snacb_last_error(None)


def test_product_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from project_morpheus_b200 import snac
    m = snac.SNAC.from_pretrained("random:0").eval().to("cpu")
    with pytest.raises(RuntimeError):
        m.decode([torch.zeros((1, k), dtype=torch.int64) for k in (1, 2, 4)])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "project_morpheus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
