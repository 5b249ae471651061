"""Import shim: the package source lives in ../intent-mpc_b200/ (a hyphen is not importable)."""
import os as _os

_real = _os.path.normpath(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), _os.pardir, "intent-mpc_b200"))
__path__ = [_real]
_init = _os.path.join(_real, "__init__.py")
exec(compile(open(_init).read(), _init, "exec"))
