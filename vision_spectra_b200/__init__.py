"""Import alias: `vision_spectra_b200` -> the package directory `vision-spectra_b200/`
(a hyphen cannot appear in a Python module name).  Sub-modules resolve inside the
real directory; this file only redirects `__path__` and runs the real `__init__`."""

import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "vision-spectra_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _os, _f
