"""Import alias: ``import uav_ac_b200`` resolves to the package directory
``uav-autonomous-control_b200/`` (whose name, fixed by the repo layout, is not a Python identifier)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "uav-autonomous-control_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
