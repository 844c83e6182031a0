"""Build recipes for the CPU oracle.  TEST INFRASTRUCTURE ONLY (see oracle/README.md).

  python oracle/build.py            # builds both targets that can be built here
  python oracle/build.py oracle     # oracle/liboracle.so from oracle/rans_oracle.c (gcc)
  python oracle/build.py ref        # oracle/_ref/rans/rans.<abi>.so from /root/reference/rans/rans.pyx

`ref` re-cythonises the reference's own rans.pyx *where it lies* under /root/reference
(its shipped rans.cpp targets CPython 3.8 and does not compile on 3.12; its own
build recipe, rans/setup.py:5-14, is `cythonize(Extension("rans", ["rans.pyx"]))`).
The generated C++ goes to a temp dir; only the compiled module lands in oracle/_ref/
(git-ignored, but it travels to the GPU box with the snapshot).  No reference source
is copied into the repo.  /root/reference does not exist on the GPU box: there the
prebuilt files are used as they are.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_PKG = os.path.join(REF_DIR, "rans")  # `from rans.rans import encode, decode` (trainer.py:32)


def _newer(target: str, *sources: str) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def build_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "rans_oracle.c")
    if not force and _newer(ORACLE_SO, src):
        return ORACLE_SO
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-pthread",
           "-o", ORACLE_SO, src, "-lm"]
    subprocess.check_call(cmd)
    return ORACLE_SO


def ref_module_path() -> str:
    return os.path.join(REF_PKG, "rans" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_ref(force: bool = False) -> str | None:
    """Compile the reference's rans.pyx into oracle/_ref/rans/.  Returns the path of the
    built module, or None when /root/reference is absent and nothing was prebuilt."""
    pyx = os.path.join(REFERENCE, "rans", "rans.pyx")
    out = ref_module_path()
    if not os.path.exists(pyx):
        return out if os.path.exists(out) else None
    if not force and _newer(out, pyx):
        return out
    os.makedirs(REF_PKG, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="flic_ref_") as tmp:
        cpp = os.path.join(tmp, "rans.cpp")
        # same translation the reference's setup.py requests (cythonize defaults, C++ from the
        # `# distutils: language=c++` header of rans.pyx), output redirected out of the read-only tree
        subprocess.check_call([sys.executable, "-m", "cython", "--cplus", pyx, "-o", cpp])
        inc = sysconfig.get_paths()["include"]
        # setuptools' default optimisation level for extension modules; baseline x86-64 (no FMA)
        cmd = ["g++", "-O2", "-fwrapv", "-fPIC", "-shared", "-I", inc, cpp, "-o", out]
        subprocess.check_call(cmd)
    return out


def main(argv: list[str]) -> int:
    what = argv[1:] or ["oracle", "ref"]
    if "oracle" in what:
        print("built", build_oracle(force="--force" in what))
    if "ref" in what:
        print("built", build_ref(force="--force" in what))
    if "clean" in what:
        shutil.rmtree(REF_DIR, ignore_errors=True)
        if os.path.exists(ORACLE_SO):
            os.remove(ORACLE_SO)
    return 0


if __name__ == "__main__":
    raise SystemExit(main(sys.argv))
