"""Importable alias for the package directory ``network-slicing_b200/`` (a hyphen is not
a legal Python identifier).  ``import ranslice_b200`` executes that package in place."""
import os as _os

_pkg = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "network-slicing_b200")
__path__ = [_pkg]
with open(_os.path.join(_pkg, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_pkg, "__init__.py"), "exec"))
del _f
