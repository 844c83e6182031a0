"""Drop-in for the reference's `rans` package: put this directory's parent on PYTHONPATH and the
reference's own import lines resolve unmodified --

    from rans.rans import encode, decode      # trainer.py:32, coder.py:15
    from rans import encode, decode           # rans/test.py:1 (run from inside rans/)

-- to the CUDA coder of this repository (flic_b200.rans: same signatures, Python lists in and
out, decode's inputs reversed by the caller as with the reference).  There is no CPU path: without
the built library or a GPU the calls raise.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from flic_b200.rans import decode, encode  # noqa: E402

__all__ = ["encode", "decode"]
