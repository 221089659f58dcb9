"""Import alias: the product package lives in the directory `track-mjx_b200/` (the name the
build contract fixes); a hyphen is not importable, so this shim points `track_mjx_b200` at it."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "track-mjx_b200")]
_init = _os.path.join(__path__[0], "__init__.py")
if _os.path.exists(_init):
    with open(_init) as _f:
        exec(compile(_f.read(), _init, "exec"))
