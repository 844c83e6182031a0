"""`rans.rans`: the module name of the reference's compiled extension (rans/setup.py:5-14 builds
`rans` inside the namespace directory `rans/`, hence `from rans.rans import ...`, trainer.py:32)."""
from . import decode, encode  # noqa: F401

__all__ = ["encode", "decode"]
