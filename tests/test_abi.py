"""The C-ABI library loads and exports every symbol include/flic_b200.h declares (no compute)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "flic_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(flic_[a-z0-9_]+)\s*\(", text)))


def test_header_and_library_agree():
    from flic_b200 import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in flic_b200.h but not exported"
    # the ctypes table binds exactly the declared functions
    assert sorted(_lib.SIGNATURES) == names
    assert L.flic_abi_version() == 2


def test_header_compiles_as_plain_c(tmp_path):
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "flic_b200.h"\nint main(void){return FLIC_ABI_VERSION - 1;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path / "t.o")])


def test_every_entry_point_cites_the_reference():
    text = open(os.path.join(ROOT, "include", "flic_b200.h")).read()
    for ref in ("rans/rans.pyx:37", "rans/rans.pyx:69", "rans/rans.pyx:49", "couplelib.py:49", "roundlib.py:18",
                "invertible.py:38", "extenddim.py:23", "trainer.py:61"):
        assert ref in text, ref


def test_argument_errors_need_no_gpu():
    """Bad arguments are rejected before any CUDA call."""
    from flic_b200 import _lib
    L = _lib.lib()
    assert L.flic_cdf_tables(None, None, None, -1, None, None, None, None) == _lib.E_ARG
    assert L.flic_couple_add_round(None, None, 1, 4, 5, 1, 1, 8, None) == _lib.E_ARG
    assert L.flic_couple_add_round(None, None, 1, 4, 3, 1, 2, 8, None) == _lib.E_ARG
    assert L.flic_squeeze(None, None, 1, 3, 5, 4, 2, 1, None) == _lib.E_ARG
    assert b"" != L.flic_last_error()
    assert L.flic_encode_workspace_bytes(1000, 10) >= 4000 + 80
