"""Importable alias for the ``neural-audio-fp_b200/`` package directory.

The package directory carries the repository's name (with hyphens), which Python cannot
import directly; this shim points its own ``__path__`` at that directory so that
``import nafp_b200.model.generate`` etc. resolve there.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "neural-audio-fp_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
