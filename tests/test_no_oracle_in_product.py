"""The oracle is test infrastructure: the product package must not import, link or run it."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "finalproject-losslessimagecompression_b200")


def test_product_never_touches_oracle():
    bad = []
    for d, _, files in os.walk(PKG):
        if "_obj" in d or "__pycache__" in d:
            continue
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                continue
            text = open(os.path.join(d, f)).read()
            for m in re.finditer(r"^\s*(from|import)\s+([\w\.]+)", text, flags=re.M):
                if m.group(2).split(".")[0] in ("oracle", "pyoracle"):
                    bad.append((f, m.group(0)))
            if "liboracle" in text or "oracle/_ref" in text or "/root/reference" in text:
                bad.append((f, "path"))
            if re.search(r'#include\s+"[^"]*oracle', text):
                bad.append((f, "include"))
    assert not bad, bad


def test_product_has_no_cpu_fallback():
    """Product entry points raise on CPU tensors instead of computing something."""
    import pytest
    import torch
    from flic_b200 import _lib, couplelib, extenddim, invertible, rans
    x = torch.zeros(4)
    with pytest.raises(TypeError):
        rans.encode_streams(x, x, x)
    with pytest.raises(TypeError):
        rans.cdf_tables(x, x, x)
    img = torch.zeros(1, 4, 2, 2)
    with pytest.raises(_lib.FlicError):
        couplelib.couple_add_round(img, img[:, 3:], 3, 1)
    from flic_b200.distlib import DLogistic
    with pytest.raises(_lib.FlicError):
        DLogistic().log_prob(img, img, img)            # only the autograd (training) path runs as torch ops
    with pytest.raises(_lib.FlicError):
        DLogistic().sample(img, img)
    with pytest.raises(_lib.FlicError):
        extenddim.squeeze(img, 2, 1)
    with pytest.raises(_lib.FlicError):
        invertible.permute_channels(img, torch.zeros(4, dtype=torch.int32))
    if not torch.cuda.is_available():
        with pytest.raises(_lib.FlicError):
            rans.HostCodec(16)
        with pytest.raises(_lib.FlicError):
            rans.encode(1 << 32, 1, [0.0], [0.0], [1.0])


def test_missing_library_fails_loudly(monkeypatch):
    import pytest
    from flic_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libflic_b200.so")
    with pytest.raises(_lib.FlicError, match="no CPU fallback"):
        _lib.lib()
