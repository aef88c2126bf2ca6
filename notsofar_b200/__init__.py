"""Import alias: the package lives in ``notsofar1-challenge_b200/`` (a name Python cannot import
directly); this shim points ``notsofar_b200`` at that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "notsofar1-challenge_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
