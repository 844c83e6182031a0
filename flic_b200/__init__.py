"""Importable alias of the product package.

The package directory is named `finalproject-losslessimagecompression_b200/` (with a hyphen, as
the project layout prescribes), which Python cannot import by name.  This stub makes
`import flic_b200` / `from flic_b200 import rans` resolve to that directory.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "finalproject-losslessimagecompression_b200")
__path__ = [_PKG_DIR]
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
del _os, _f
