"""Importable alias for the package directory ``visuelle2-multimodal-fusion_b200/``.

The product lives in the hyphenated directory the build contract names; a hyphen cannot appear
in a Python import, so this stub points ``__path__`` at that directory and runs its ``__init__``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "visuelle2-multimodal-fusion_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f
